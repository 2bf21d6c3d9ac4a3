#!/bin/bash
# Round 2, GPU call Q: L2 replacement hints in the tail kernel (RVCR_RC_CS 0 / 1 / 2, see scan_tc.cu).
set -u
mkdir -p gpurun_out
PKG=codename-rvc-fork-3_b200
cp $PKG/librvcr.so /tmp/librvcr_default.so
run() {
  local name=$1
  ( timeout 200 python bench.py --steps 3 --no-cpu-baseline --no-e2e --query-skew 0 ) 2>gpurun_out/r02q_bench_$name.err | grep '^{' > gpurun_out/r02q_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02q_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j["phase_ms"].items() if k.startswith("scan")})
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02q_bench_{sys.argv[1]}.err").read()[-1200:])
PY
}
run cs0
for v in cs1 cs2; do
  cp variants/librvcr_$v.so $PKG/librvcr.so
  ( timeout 120 python -m pytest tests/test_gpu_scan_tc.py -x -q ) > gpurun_out/r02q_pytest_$v.log 2>&1; echo "$v scan tests rc=$?"
  run $v
done
cp /tmp/librvcr_default.so $PKG/librvcr.so
