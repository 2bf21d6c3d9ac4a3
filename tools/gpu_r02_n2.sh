#!/bin/bash
# Round 2, 2-GPU call: the device-driven sharded step (inboxes + flag barriers) -- parity test, then the bench at N = 2.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r02_n2.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
( time timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q -s --timeout 350 ) > gpurun_out/r02n2b_pytest_sharded.log 2>&1
rc=$?; echo "sharded rc=$rc" >> gpurun_out/r02n2b_pytest_sharded.log; tail -15 gpurun_out/r02n2b_pytest_sharded.log
if [ $rc -ne 0 ]; then exit 1; fi
( NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r02n2b_bench.log 2> gpurun_out/r02n2b_bench.err
grep '^{' gpurun_out/r02n2b_bench.log > gpurun_out/r02n2b_bench.json
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r02n2b_bench.json"))
    print("N=2", round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()}, "e2e", j["e2e"]["ms_per_step"],
          "parity", j["config"].get("parity_sample"), "skew", j.get("skewed_queries"), j["config"].get("transport"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r02n2b_bench.err").read()[-2500:])
PY
( RVCR_PEER_PULL=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 ) 2>/dev/null | grep '^{' | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('N=2 pull-first variant', round(j['ms_per_step'],2), {k: round(v,2) for k,v in j['phase_ms'].items()})"
