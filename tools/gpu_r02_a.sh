#!/bin/bash
# Round 2, GPU call A: everything that has never run on a B200 gets its first run.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02_a.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_smi.txt 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
( timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -1
( RVCR_TEST_SCAN_IMPL=6 timeout 300 python -m pytest tests/test_gpu_scan_tc.py -x -q ) > gpurun_out/r02a_pytest_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/r02a_pytest_fused.log
tail -6 gpurun_out/r02a_pytest_fused.log
( RVCR_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_experimental.py -q -s ) > gpurun_out/r02a_pytest_experimental.log 2>&1
echo "experimental rc=$?" >> gpurun_out/r02a_pytest_experimental.log
tail -12 gpurun_out/r02a_pytest_experimental.log
run() {  # name, env..., -- bench args
  local name=$1; shift
  ( env "$@" timeout 400 python bench.py --steps 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} ) 2>gpurun_out/r02a_bench_$name.err | grep '^{' > gpurun_out/r02a_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02a_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j["phase_ms"].items()}, j["scan_stats"], j.get("coarse_stats"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default RVCR_NOP=1
BENCH_ARGS="--scan-impl 6" run fused RVCR_NOP=1
BENCH_ARGS="--coarse-mode 4" run coarse_e4m3 RVCR_NOP=1
BENCH_ARGS="--coarse-mode 4 --scan-impl 6" run coarse_e4m3_fused RVCR_NOP=1
run rc_8_16 RVCR_RC_CTAS=8 RVCR_RC_UNROLL=16
run rc_10_16 RVCR_RC_CTAS=10 RVCR_RC_UNROLL=16
