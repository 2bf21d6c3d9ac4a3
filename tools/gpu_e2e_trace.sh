#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/e2e_*.log
RVCR_HOST_TRACE=1 timeout 600 python bench.py --steps 1 --no-cpu-baseline > gpurun_out/e2e_trace.log 2>&1
grep -E "trace" gpurun_out/e2e_trace.log | tail -25
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e_*.log')):
    for ln in open(f):
        if ln.startswith('{'):
            j=json.loads(ln); print(f, round(j['value']), round(j['ms_per_step'],2), round(j['e2e']['value']), round(j['e2e']['ms_per_step'],2))
PY
