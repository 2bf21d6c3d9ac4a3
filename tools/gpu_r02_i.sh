#!/bin/bash
# Round 2, GPU call I: coarse slow path per group (8 insertion sites), tail-kernel occupancy x unroll sweep.
set -u
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_coarse_tc.py tests/test_gpu_coarse_f8.py -x -q ) > gpurun_out/r02i_pytest_coarse.log 2>&1
rc=$?; echo "coarse rc=$rc" >> gpurun_out/r02i_pytest_coarse.log; tail -4 gpurun_out/r02i_pytest_coarse.log
if [ $rc -ne 0 ]; then echo "coarse tests failed: stopping"; tail -40 gpurun_out/r02i_pytest_coarse.log; exit 1; fi
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 python bench.py "$@" ) 2>gpurun_out/r02i_bench_$name.err | grep '^{' > gpurun_out/r02i_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02i_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()}, "clk", j.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02i_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
B="--steps 3 --no-cpu-baseline --no-e2e --query-skew 0"
run e4m3 X=1 -- $B
run bf16 X=1 -- $B --coarse-f8-auto 0
for c in 2 3 4 6; do run rc_${c}_16 RVCR_RC_CTAS=$c RVCR_RC_UNROLL=16 -- $B; done
run rc_4_8 RVCR_RC_CTAS=4 -- $B
run rc_8_16 RVCR_RC_CTAS=8 RVCR_RC_UNROLL=16 -- $B
