#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        j=json.loads(ln); p=j['phase_ms']; print(j['value'], j['ms_per_step'], j['roofline']['achieved'], j['roofline']['frac'], j.get('scan_stats'), j['clocks']); print(p)
    else: print(ln[:200])
"
