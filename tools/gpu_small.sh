#!/bin/bash
# bench lines of the small / medium configs (C1, C2, C3) for profiles/
set -u
mkdir -p gpurun_out
for w in C1 C2 C3; do
  ( time timeout 600 python bench.py --workload $w --steps 20 ) > gpurun_out/bench_${w}_n1.log 2>&1
  grep '^{' gpurun_out/bench_${w}_n1.log | python -c "
import sys,json
for ln in sys.stdin:
    j=json.loads(ln); print('$w', round(j['value']), round(j['ms_per_step'],4), j['phase_ms'], j['e2e'], j['cpu_baseline'], j['gpu_launches'])
"
  tail -2 gpurun_out/bench_${w}_n1.log | cut -c1-300
done
