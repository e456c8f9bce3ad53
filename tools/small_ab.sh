#!/bin/bash
# GPU box: A/B of the small-net thread mappings (lane = scenario; lane = hidden unit with 1 / 2 / 4 scenarios per warp)
# on cfg 1 / cfg 3 (+ lead 20) at training batch sizes: device-timed forward + adjoint step
timeout 900 python -m pytest tests/test_kernels_abi.py tests/test_full_size_properties.py -m gpu -x -q -k "unit or small or one_store or serial" > gpurun_out/r2g_small_tests.log 2>&1
tail -3 gpurun_out/r2g_small_tests.log
out=gpurun_out/r2g_small_ab.log; rm -f $out
run() { env $1 $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-others --workload "${@:3}" 2>&1 | python -c "
import sys, json
tag=' '.join(sys.argv[1:])
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'):
        continue
    d=json.loads(line)
    print('%-60s ms_per_step %.4f  value %.1fM' % (tag, d['ms_per_step'], d['value']/1e6))
" "$@" >> $out; }
for wl in "one_store_lost --batch 1024" "one_store_lost --batch 2048" "one_store_lost --batch 4096" "one_store_lost --batch 8192" "one_store_lost --batch 16384" "one_store_lost --batch 32768" "serial_system --batch 2048" "serial_system --batch 4096" "serial_system --batch 8192" "serial_system --batch 16384" "one_store_backlogged_lead20 --batch 8192"; do
  run HDPO_SMALL_UNIT_MAX=0 HDPO_SMALL_UNIT_G=0 $wl
  for g in 1 2 4; do run HDPO_SMALL_UNIT_MAX=1000000 HDPO_SMALL_UNIT_G=$g $wl; done
done
cat $out
