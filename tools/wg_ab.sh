out=gpurun_out/r2v_ksplit_n.log; rm -f $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
for rep in 1 2; do
WL=one_warehouse_lost_demand
run HDPO_X=default
run HDPO_WIDE_KSPLIT=1 HDPO_WIDE_KSPLIT_N=2
WL=many_warehouses_lost_demand
run HDPO_X=default
run HDPO_WIDE_KSPLIT_N=2
run HDPO_WIDE_KSPLIT_N=8
done
cat $out
