#!/bin/bash
# Round 2, GPU call L: streaming scan with ONE query operand buffer (RVCR_STC_NQB=1: 9 row stages in flight instead of 6), with and
# without 2-tile chunks (10 stages).  Call K showed the kernel is bound by bytes in flight: 3 / 4 / 5 / 6 stages -> 35.7 / 28.8 / 23.8 / 19.7 ms.
set -u
mkdir -p gpurun_out
PKG=codename-rvc-fork-3_b200
cp $PKG/librvcr.so /tmp/librvcr_default.so
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 python bench.py "$@" ) 2>gpurun_out/r02l_bench_$name.err | grep '^{' > gpurun_out/r02l_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02l_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()}, "clk", (j.get("clocks") or {}).get("sm_mhz"), (j.get("config") or {}).get("parity_sample"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02l_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
B="--steps 3 --no-e2e --query-skew 0"
for v in nqb1 nqb1c2; do
  cp variants/librvcr_$v.so $PKG/librvcr.so
  ( timeout 200 python -m pytest tests/test_gpu_scan_tc.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r02l_pytest_$v.log 2>&1
  rc=$?; echo "$v tests rc=$rc"; tail -3 gpurun_out/r02l_pytest_$v.log
  if [ $rc -ne 0 ]; then tail -40 gpurun_out/r02l_pytest_$v.log; continue; fi
  run $v X=1 -- $B
  run ${v}_C3 X=1 -- $B --workload C3
done
cp /tmp/librvcr_default.so $PKG/librvcr.so
run default X=1 -- $B --no-cpu-baseline
