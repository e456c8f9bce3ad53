# GPU box: full GPU suite, smoke(), the default bench line (with other_workloads, cpu_baseline, reference_cuda), reference arm
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_gpu_tests.log 2>&1; tail -2 gpurun_out/r2l_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; tail -1 gpurun_out/r2l_smoke.log
timeout 1500 python bench.py > gpurun_out/r2l_bench_default.json 2> gpurun_out/r2l_bench_default.err; tail -c 300 gpurun_out/r2l_bench_default.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.err; tail -c 400 gpurun_out/r2l_bench_reference.json
