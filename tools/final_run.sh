# GPU box: full GPU suite, smoke(), K3 A/B, the default bench line (with other_workloads, cpu_baseline, reference_cuda)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gpu_tests.log 2>&1; tail -2 gpurun_out/r2h_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
python tools/step_time.py > gpurun_out/r2h_step_time.log 2>&1
HDPO_STEP_THREAD=1 python tools/step_time.py >> gpurun_out/r2h_step_time.log 2>&1
B=1024 python tools/step_time.py >> gpurun_out/r2h_step_time.log 2>&1
B=1024 HDPO_STEP_THREAD=1 python tools/step_time.py >> gpurun_out/r2h_step_time.log 2>&1
cat gpurun_out/r2h_step_time.log
timeout 1500 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; tail -c 600 gpurun_out/r2h_bench_default.json
