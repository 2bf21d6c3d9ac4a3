#!/bin/bash
# 2-GPU box: sharded parity test + the bench at N=2 (and N=1 beside it for the ratio)
set -u
mkdir -p gpurun_out
N=${1:-2}
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/pytest_sharded.log 2>&1
tail -4 gpurun_out/pytest_sharded.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/bench_c5_n$N.log 2>&1
grep '^{' gpurun_out/bench_c5_n$N.log | python -c "
import sys,json
for ln in sys.stdin:
    j=json.loads(ln); print(j['n_gpus'], round(j['value']), round(j['ms_per_step'],2), j['phase_ms'], j['e2e'], j['clocks'])
"
tail -3 gpurun_out/bench_c5_n$N.log | cut -c1-300
