timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gpu_tests.log 2>&1
tail -3 gpurun_out/r2f_gpu_tests.log
rm -f gpurun_out/r2f_wide_ab.log
run() { env "$@" timeout 120 python tools/wide_ab.py $WL 2>&1 | tail -1 >> gpurun_out/r2f_wide_ab.log; }
WL=one_warehouse_lost_demand
run HDPO_X=default
for ch in 3 4; do for g in 2 5 10; do run HDPO_WIDE_WG_OVERLAP=1 HDPO_WIDE_WG_GROUP=$g HDPO_WIDE_CHUNKS=$ch; done; done
WL=many_warehouses_lost_demand
run HDPO_X=default
run HDPO_WIDE_WG_OVERLAP=0
run HDPO_WIDE_WG_GROUP=2
run HDPO_WIDE_WG_GROUP=10
WL=many_warehouses_lost_demand_8192
run HDPO_X=default
cat gpurun_out/r2f_wide_ab.log
