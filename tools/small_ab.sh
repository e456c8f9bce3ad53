#!/bin/bash
# A/B of the two small-net mappings on cfg 1 / cfg 3 at the reference's train batch (device-timed step)
run() { echo "== $*"; env $1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-others --workload "${@:2}" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'):
        print(line[:300]); continue
    d=json.loads(line)
    print('ms_per_step %.4f  value %.1fM  launches/step %.0f' % (d['ms_per_step'], d['value']/1e6, d['gpu_launches']/d['steps']))
"; }
for wl in "one_store_lost --batch 1024" "one_store_lost --batch 2048" "one_store_lost --batch 4096" "one_store_lost --batch 6144" "serial_system --batch 1024" "serial_system --batch 2048" "serial_system --batch 4096"; do
  run HDPO_SMALL_UNIT_MAX=0 $wl
  run HDPO_SMALL_UNIT_MAX=16384 $wl
done
