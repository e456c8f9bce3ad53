out=gpurun_out/r2t_wg_pair2.log; rm -f $out
for m in 1 0; do HDPO_WG_PAIR=$m timeout 300 python -m pytest tests/test_gemm_tc.py -m gpu -x -q -s -k weight_gradient 2>&1 | grep -E "gemm_tc wgrad|passed|failed|rror" | head -14 >> $out; done
timeout 900 python -m pytest tests/test_kernels_abi.py tests/test_full_size_properties.py tests/test_wide_variants.py tests/test_symmetry_aware_fused.py tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
for rep in 1 2 3; do
WL=one_warehouse_lost_demand
run HDPO_WG_PAIR=0
run HDPO_WG_PAIR=2
run HDPO_WG_PAIR=1
done
for WL in many_warehouses_lost_demand many_warehouses_lost_demand_8192 one_warehouse_lost_demand_symmetry_aware; do
run HDPO_WG_PAIR=0
run HDPO_WG_PAIR=1
done
cat $out
