#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/e2e_*.log
timeout 600 python -m pytest tests/test_gpu_scan_tc.py -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --no-cpu-baseline > gpurun_out/e2e_$name.log 2>&1; }
run plain RVCR_HOST_BULK_STORE=0
run plain_c256 RVCR_HOST_BULK_STORE=0 RVCR_HOST_BLEND_CTAS=256
run b2_s5_c148 RVCR_HOST_BLEND_CTAS=148
run b2_s5_c296 RVCR_HOST_BLEND_CTAS=296
run b2_s4_c592 RVCR_HOST_BLEND_CTAS=592 RVCR_STC_NSTAGE=4
run b2_s4_c296 RVCR_HOST_BLEND_CTAS=296 RVCR_STC_NSTAGE=4
run b4_s4_c296 RVCR_HOST_BLEND_CTAS=296 RVCR_STC_NSTAGE=4 RVCR_HOST_NBUF=4
run b2_s3_c888 RVCR_HOST_BLEND_CTAS=888 RVCR_STC_NSTAGE=3
RVCR_HOST_TRACE=1 RVCR_HOST_BLEND_CTAS=296 timeout 600 python bench.py --steps 2 --no-cpu-baseline > gpurun_out/e2e_trace.log 2>&1
grep -E "trace" gpurun_out/e2e_trace.log | tail -16
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e_*.log')):
    ok=False
    for ln in open(f):
        if ln.startswith('{'):
            ok=True
            j=json.loads(ln); print(f, round(j['value']), round(j['ms_per_step'],2), round(j['e2e']['value']), round(j['e2e']['ms_per_step'],2))
    if not ok: print(f, open(f).read()[-800:])
PY
