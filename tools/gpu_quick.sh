#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py smoke ) 2>&1 | tail -1
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or config_parity" 2>&1 | tail -2
timeout 100 python bench.py --workload C1 --steps 30 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_C1_auto_v2.json
python -c "
import json
j=json.load(open('gpurun_out/bench_C1_auto_v2.json')); print('C1 auto', round(j['ms_per_step']*1000,1), 'us', {k:round(v*1000,1) for k,v in j['phase_ms'].items()})"
