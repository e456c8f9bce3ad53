# GPU box: parity of everything + timings after the packed-f32x2 epilogue / ELU
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gpu_tests.log 2>&1; tail -2 gpurun_out/r2i_gpu_tests.log
rm -f gpurun_out/r2i_ab.log
for wl in one_warehouse_lost_demand one_warehouse_lost_demand_symmetry_aware many_warehouses_lost_demand; do
  timeout 200 python tools/wide_ab.py $wl 2>&1 | tail -1 >> gpurun_out/r2i_ab.log
done
for wl in "one_store_lost" "one_store_backlogged_lead20" "serial_system"; do
 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-others --workload $wl 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print(d['config']['workload'], 'ms_per_step %.3f' % d['ms_per_step'])
" >> gpurun_out/r2i_ab.log
done
cat gpurun_out/r2i_ab.log
