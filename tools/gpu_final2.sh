#!/bin/bash
# Rehearsal after the auto-dispatch change: full GPU suite (-x), smoke(), C1 / C2 latency with the size rules (auto)
# against the forced tensor-core kernels, and the C5 headline line.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_final2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final2.log
tail -5 gpurun_out/pytest_final2.log
( timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -1
for w in C1 C2; do
  for mode in "--coarse-mode 0 --scan-impl 0" "--coarse-mode 2 --scan-impl 5"; do
    timeout 300 python bench.py --workload $w --steps 30 --no-cpu-baseline $mode 2>/dev/null | grep '^{' > gpurun_out/bench_${w}_$(echo $mode | tr -d ' -').json
    python -c "
import sys,json
j=json.load(open('gpurun_out/bench_${w}_$(echo $mode | tr -d ' -').json')); print('$w', '$mode', round(j['ms_per_step']*1000,1), 'us', round(j['e2e']['ms_per_step']*1000,1), 'us e2e', {k:round(v*1000,1) for k,v in j['phase_ms'].items()})"
  done
done
timeout 600 python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_c5_final2.json
python -c "
import json
j=json.load(open('gpurun_out/bench_c5_final2.json')); print('C5', round(j['value']), round(j['ms_per_step'],2), round(j['e2e']['value']), j['phase_ms'])"
