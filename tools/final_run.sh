# GPU box: full GPU suite, smoke(), wide workloads fwd/bwd, the default bench line, reference arm
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_gpu_tests.log 2>&1; tail -2 gpurun_out/r2o_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; tail -1 gpurun_out/r2o_smoke.log
for wl in one_warehouse_lost_demand many_warehouses_lost_demand many_warehouses_lost_demand_8192 one_warehouse_lost_demand_symmetry_aware; do python tools/wide_ab.py $wl | tail -1; done
HDPO_AB_BATCH=1024 python tools/wide_ab.py one_warehouse_lost_demand | tail -1
timeout 1500 python bench.py > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err; tail -c 200 gpurun_out/r2o_bench_default.json
