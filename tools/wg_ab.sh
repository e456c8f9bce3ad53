out=gpurun_out/r2p_prio_ab2.log; rm -f $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
for rep in 1 2; do
WL=one_warehouse_lost_demand
run HDPO_X=default
for g in 2 5 10; do
run HDPO_WIDE_WG_OVERLAP=1 HDPO_WIDE_PRIO=3 HDPO_WIDE_WG_GROUP=$g
done
run HDPO_WIDE_WG_OVERLAP=1 HDPO_WIDE_PRIO=0 HDPO_WIDE_WG_GROUP=5
WL=many_warehouses_lost_demand
run HDPO_X=default
run HDPO_WIDE_PRIO=3
done
cat $out
