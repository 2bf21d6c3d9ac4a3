#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_scan_tc.py -x -q -s ) > gpurun_out/pytest_scan_tc.log 2>&1
grep -E "scan_tc:|passed|failed|Error" gpurun_out/pytest_scan_tc.log | head
for f in 0; do
  echo "== RVCR_STC_FLAGS=$f"
  RVCR_STC_FLAGS=$f timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        j=json.loads(ln); p=j['phase_ms']; print(j['value'], j['ms_per_step'], 'scan', p['scan'], 'recheck', p.get('scan_recheck'), 'tc', p['scan']-p.get('scan_recheck',0)-p.get('scan_fallback',0), j.get('scan_stats')); print(p)
    else: print(ln[:200])
"
done
( timeout 600 python tools/perf_probe.py --mode 2 --scan 5 ) > gpurun_out/probe_s5.log 2>&1
tail -2 gpurun_out/probe_s5.log | head -1
