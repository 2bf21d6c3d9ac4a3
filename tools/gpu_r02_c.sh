#!/bin/bash
# Round 2, GPU call C: coarse epilogue v3 (candidates within the margin of the running best instead of a sorted top-8),
# nprobe > 1 search, promoted e4m3 tests.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_coarse_tc.py tests/test_gpu_coarse_f8.py tests/test_gpu_nprobe.py tests/test_gpu_interop.py tests/test_gpu_parity.py -x -q -s ) > gpurun_out/r02c_pytest_a.log 2>&1
echo "rc=$?" >> gpurun_out/r02c_pytest_a.log; grep -E "e4m3|passed|failed|rc=|Error|assert" gpurun_out/r02c_pytest_a.log | tail -15
run() {  # name, bench args
  local name=$1; shift
  ( timeout 600 python bench.py "$@" ) 2>gpurun_out/r02c_bench_$name.err | grep '^{' > gpurun_out/r02c_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02c_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()},
          j.get("coarse_stats"), "parity", j["config"].get("parity_sample"), "clk", j.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02c_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
run c5_bf16 --steps 3 --no-e2e --query-skew 0
run c5_e4m3 --steps 3 --coarse-mode 4 --no-e2e --query-skew 0
run c5_f8auto --steps 3 --coarse-f8-auto 1 --no-cpu-baseline --no-e2e --query-skew 0
run c5_e4m3_cl1 --steps 3 --coarse-mode 4 --no-cpu-baseline --no-e2e --query-skew 0 --tc-cluster 1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel' -c 1 -o gpurun_out/r02c_coarse_e4m3 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 --coarse-mode 4 ) > gpurun_out/r02c_ncu_e4m3.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel' -c 1 -o gpurun_out/r02c_coarse_bf16 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 ) > gpurun_out/r02c_ncu_bf16.log 2>&1
ls -la gpurun_out | tail -5
