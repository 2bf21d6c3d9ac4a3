#!/bin/bash
# Round 2, GPU call G: coarse kernel back on 2 x 256 accumulators, per-row floor shared by the column-group warps.
set -u
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_coarse_tc.py tests/test_gpu_coarse_f8.py -x -q ) > gpurun_out/r02h_pytest_coarse.log 2>&1
rc=$?; echo "coarse rc=$rc" >> gpurun_out/r02h_pytest_coarse.log; tail -4 gpurun_out/r02h_pytest_coarse.log
if [ $rc -ne 0 ]; then echo "coarse tests failed: stopping"; tail -40 gpurun_out/r02h_pytest_coarse.log; exit 1; fi
run() {  # name, bench args
  local name=$1; shift
  ( timeout 300 python bench.py "$@" ) 2>gpurun_out/r02h_bench_$name.err | grep '^{' > gpurun_out/r02h_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02h_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()},
          j.get("coarse_stats"), "parity", j["config"].get("parity_sample"), "clk", j.get("clocks", {}).get("sm_mhz"),
          "e2e", j.get("e2e") and round(j["e2e"]["ms_per_step"], 1))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02h_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
run c5_bf16 --steps 3 --no-cpu-baseline --no-e2e --query-skew 0
run c5_e4m3 --steps 3 --coarse-mode 4 --no-cpu-baseline --no-e2e --query-skew 0
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel' -c 1 -o gpurun_out/r02h_coarse_e4m3 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 --coarse-mode 4 ) > gpurun_out/r02h_ncu_e4m3.log 2>&1
