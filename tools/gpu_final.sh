#!/bin/bash
# Round-end rehearsal on one GPU: full GPU suite (-x like the driver), smoke(), the default bench line, and two
# bench runs with the re-check occupancy knob.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -6 gpurun_out/pytest_final.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_final.log 2>&1
grep '^{' gpurun_out/bench_final.log | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],2), j['roofline']['traffic'], round(j['e2e']['value']), j['phase_ms'])"
for c in 10 6; do
  RVCR_RC_CTAS=$c timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('RC_CTAS=$c', round(j['value']), round(j['ms_per_step'],2), round(j['phase_ms']['scan'],2), round(j['phase_ms']['scan_recheck'],2))"
done
