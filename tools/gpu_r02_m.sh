#!/bin/bash
# Round 2, GPU call M: L2 prefetch ahead of the register / shared-memory landing buffers.  RVCR_STC_PREFETCH = TMA L2 prefetch n
# stages ahead of the scan kernel's row ring; RVCR_RC_PREFETCH = the tail kernel prefetches every candidate row to L2 up front.
set -u
mkdir -p gpurun_out
( time RVCR_STC_PREFETCH=8 RVCR_RC_PREFETCH=1 timeout 200 python -m pytest tests/test_gpu_scan_tc.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r02m_pytest.log 2>&1
rc=$?; echo "tests rc=$rc" >> gpurun_out/r02m_pytest.log; tail -4 gpurun_out/r02m_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed: stopping"; tail -40 gpurun_out/r02m_pytest.log; exit 1; fi
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 python bench.py "$@" ) 2>gpurun_out/r02m_bench_$name.err | grep '^{' > gpurun_out/r02m_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02m_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items() if k.startswith("scan")}, "clk", (j.get("clocks") or {}).get("sm_mhz"), (j.get("config") or {}).get("parity_sample"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02m_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
B="--steps 3 --no-cpu-baseline --no-e2e --query-skew 0"
run base X=1 -- $B
for p in 3 6 12 24 48; do run stcpf$p RVCR_STC_PREFETCH=$p -- $B; done
for p in 1 5; do run rcpf$p RVCR_RC_PREFETCH=$p -- $B; done
for c in 4 6; do run rcpf1_ctas$c RVCR_RC_PREFETCH=1 RVCR_RC_CTAS=$c -- $B; done
run both RVCR_STC_PREFETCH=12 RVCR_RC_PREFETCH=1 -- --steps 3 --no-e2e --query-skew 0
