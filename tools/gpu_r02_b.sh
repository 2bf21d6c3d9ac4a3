#!/bin/bash
# Round 2, GPU call B: the re-structured coarse epilogue (16 epilogue warps, norms by bulk copy, pipelined TMEM loads),
# the re-worked bench (ring buffers, parity sample at every N, skewed queries, C4 workload) and smoke().
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_coarse_tc.py -x -q ) > gpurun_out/r02b_pytest_coarse.log 2>&1
echo "coarse rc=$?" >> gpurun_out/r02b_pytest_coarse.log; tail -4 gpurun_out/r02b_pytest_coarse.log
( RVCR_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_experimental.py -q -s -k "f8" ) > gpurun_out/r02b_pytest_f8.log 2>&1
echo "f8 rc=$?" >> gpurun_out/r02b_pytest_f8.log; grep -E "e4m3|passed|failed|rc=" gpurun_out/r02b_pytest_f8.log | tail -8
( timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -2
run() {  # name, bench args
  local name=$1; shift
  ( timeout 600 python bench.py "$@" ) 2>gpurun_out/r02b_bench_$name.err | grep '^{' > gpurun_out/r02b_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02b_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()},
          j.get("coarse_stats"), j.get("scan_stats"), "e2e", j.get("e2e") and round(j["e2e"]["ms_per_step"], 1),
          "parity", j["config"].get("parity_sample"), "skew", j.get("skewed_queries"), "clk", j.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02b_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
run c5_default --steps 5
run c5_e4m3 --steps 3 --coarse-mode 4 --no-cpu-baseline --no-e2e --query-skew 0
run c5_f8auto --steps 3 --coarse-f8-auto 1 --no-cpu-baseline --no-e2e --query-skew 0
run c4 --workload C4 --steps 3
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log; tail -5 gpurun_out/r02b_pytest.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel' -c 1 -o gpurun_out/r02b_coarse_bf16 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 ) > gpurun_out/r02b_ncu_bf16.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel' -c 1 -o gpurun_out/r02b_coarse_e4m3 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 --coarse-mode 4 ) > gpurun_out/r02b_ncu_e4m3.log 2>&1
ls -la gpurun_out | tail -8
