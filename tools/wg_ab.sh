out=gpurun_out/r2q_ksplit.log; rm -f $out
timeout 900 python -m pytest tests/test_kernels_abi.py tests/test_full_size_properties.py tests/test_trainer_gpu.py tests/test_wide_variants.py tests/test_wide_persistent.py tests/test_gemm_tc.py tests/test_gemm_multi.py -m gpu -x -q -k "wide or warehouse or gemm or persist or variant" 2>&1 | tail -3 >> $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
for rep in 1 2 3; do
WL=one_warehouse_lost_demand
run HDPO_WIDE_KSPLIT=1
run HDPO_WIDE_KSPLIT=0
done
WL=many_warehouses_lost_demand
run HDPO_WIDE_KSPLIT=1
run HDPO_WIDE_KSPLIT=0
WL=many_warehouses_lost_demand_8192
run HDPO_WIDE_KSPLIT=1
run HDPO_WIDE_KSPLIT=0
WL=one_warehouse_lost_demand
run HDPO_AB_BATCH=1024 HDPO_WIDE_KSPLIT=1
run HDPO_AB_BATCH=1024 HDPO_WIDE_KSPLIT=0
cat $out
