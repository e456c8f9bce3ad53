#!/bin/bash
# A/B of the persistent wide path on the default bench workload (device-timed step, forward, adjoint)
run() { echo "== $1"; env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-others ${2:-} 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'):
        print(line[:300]); continue
    d=json.loads(line); r=d['roofline']
    print('ms_per_step %.3f  fwd %.3f  bwd %.3f  launches/step %.0f  value %.2fM' % (d['ms_per_step'], r['fwd_kernel_ms'], r.get('adjoint_group',r)['kernel_ms'], d['gpu_launches']/d['steps'], d['value']/1e6))
"; }
for cfg in "$@"; do run "$cfg"; done
