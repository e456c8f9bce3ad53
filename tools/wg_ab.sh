rm -f gpurun_out/r2k_persist_small.log
run() { env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 >> gpurun_out/r2k_persist_small.log; }
WL=many_warehouses_lost_demand
run HDPO_X=default
run HDPO_WIDE_PERSIST=1
run HDPO_WIDE_PERSIST=1 HDPO_WIDE_PERSIST_BWD=0
WL=one_warehouse_lost_demand
for b in 1024 2048; do
run HDPO_AB_BATCH=$b
run HDPO_AB_BATCH=$b HDPO_WIDE_PERSIST=1
run HDPO_AB_BATCH=$b HDPO_WIDE_PERSIST=1 HDPO_WIDE_PERSIST_BWD=0
run HDPO_AB_BATCH=$b HDPO_TC_MULTI=1
done
cat gpurun_out/r2k_persist_small.log
