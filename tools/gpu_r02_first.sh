#!/bin/bash
# First GPU call of round 2 (everything written after the last GPU session of round 1 gets its first run here):
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02_first.sh'
# 1. the full GPU suite (now with the legacy-layout reader, tests/test_gpu_zz_legacy_index.py) and smoke()
# 2. the experimental fused scan + re-check kernel (scan_impl 6) against the same tests as scan_impl 5 -- under its own
#    timeout; its mbarrier waits are bounded (trap, not hang)
#    and the e4m3 coarse step (coarse_mode 4) against the exact kernel (tests/test_gpu_experimental.py)
# 3. C5 bench lines: default, scan_impl 6, coarse_mode 4, both, and the re-check tuning knobs (CTAs per SM x unroll)
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -4 gpurun_out/r02_pytest.log
( timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -1
( RVCR_TEST_SCAN_IMPL=6 timeout 300 python -m pytest tests/test_gpu_scan_tc.py -x -q ) > gpurun_out/r02_pytest_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/r02_pytest_fused.log
tail -6 gpurun_out/r02_pytest_fused.log
( RVCR_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q -s ) > gpurun_out/r02_pytest_experimental.log 2>&1
echo "experimental rc=$?" >> gpurun_out/r02_pytest_experimental.log
tail -8 gpurun_out/r02_pytest_experimental.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
run() {  # name, env..., -- bench args
  local name=$1; shift
  ( env "$@" timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} ) 2>gpurun_out/r02_bench_$name.err | grep '^{' > gpurun_out/r02_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j["phase_ms"].items()}, j["scan_stats"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default RVCR_NOP=1
BENCH_ARGS="--scan-impl 6" run fused RVCR_NOP=1
BENCH_ARGS="--scan-impl 6" run fused_paired_kblocks RVCR_STC_FLAGS=8
BENCH_ARGS="--coarse-mode 4" run coarse_e4m3 RVCR_NOP=1
BENCH_ARGS="--coarse-mode 4 --scan-impl 6" run coarse_e4m3_fused RVCR_NOP=1
BENCH_ARGS="--coarse-f8-auto 1" run coarse_auto_probe RVCR_NOP=1
for ctas in 4 6 8 10; do
  for unr in 8 16; do
    run rc_${ctas}_${unr} RVCR_RC_CTAS=$ctas RVCR_RC_UNROLL=$unr
  done
done
timeout 600 python bench.py > gpurun_out/r02_bench_c5_n1.log 2>&1; tail -c 600 gpurun_out/r02_bench_c5_n1.log
