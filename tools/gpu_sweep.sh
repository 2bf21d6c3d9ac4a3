#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/sweep_*.log
for c in 3 4 6 8 12 16; do
  RVCR_RC_CTAS=$c timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-e2e --tc-resident 9 --fused-blend 1 --recheck-impl 0 > gpurun_out/sweep_f1_c$c.log 2>&1
done
for c in 4 12 16; do
  RVCR_RC_CTAS=$c timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-e2e --tc-resident 9 --fused-blend 0 --recheck-impl 0 > gpurun_out/sweep_f0_c$c.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_*.log')):
    ok=False
    for ln in open(f):
        if ln.startswith('{'):
            ok=True
            j=json.loads(ln); print(f, round(j['value']), round(j['ms_per_step'],2), {k:round(v,2) for k,v in j['phase_ms'].items()}, j['clocks'])
    if not ok: print(f, open(f).read()[-600:])
PY
