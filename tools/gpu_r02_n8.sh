#!/bin/bash
# Round 2, 8-GPU call: sharded parity at world 8, bench at N = 8 and N = 4 (device-driven step).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r02_n8.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
( time timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -s --timeout 280 ) > gpurun_out/r02n8_pytest_sharded.log 2>&1
rc=$?; echo "sharded rc=$rc" >> gpurun_out/r02n8_pytest_sharded.log; tail -6 gpurun_out/r02n8_pytest_sharded.log
for n in 8 4; do
  ( NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 ) > gpurun_out/r02n${n}_bench.log 2> gpurun_out/r02n${n}_bench.err
  grep '^{' gpurun_out/r02n${n}_bench.log > gpurun_out/r02n${n}_bench.json
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/r02n{n}_bench.json"))
    print("N=" + n, round(j["ms_per_step"], 2), "ms/step", round(j["value"] / 1e6, 2), "M frames/s", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()},
          "e2e ms", round(j["e2e"]["ms_per_step"], 1), "parity", j["config"].get("parity_sample"), "skew", j.get("skewed_queries"))
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/r02n{n}_bench.err").read()[-2500:])
PY
done
grep -c "NCCL INFO" gpurun_out/r02n8_bench.log gpurun_out/r02n8_bench.err | head -3
