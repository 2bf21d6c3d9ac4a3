#!/bin/bash
# Round 2, GPU call N: e4m3 epilogue waits parked (mbarrier.try_wait with a suspend hint) instead of polled -- the tmem_full
# spin loop is 36 % of the kernel's executed instructions and the kernel runs under sw_power_cap; + the de-flaked drop-in test.
set -u
mkdir -p gpurun_out
PKG=codename-rvc-fork-3_b200
cp $PKG/librvcr.so /tmp/librvcr_default.so
( timeout 300 python -m pytest tests/test_gpu_dropin.py -x -q ) > gpurun_out/r02n_pytest_dropin.log 2>&1; echo "dropin rc=$?"; tail -3 gpurun_out/r02n_pytest_dropin.log
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 python bench.py "$@" ) 2>gpurun_out/r02n_bench_$name.err | grep '^{' > gpurun_out/r02n_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02n_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items() if k.startswith("coarse")}, "clk", (j.get("clocks") or {}).get("sm_mhz"), (j.get("config") or {}).get("parity_sample"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02n_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
B="--steps 5 --no-cpu-baseline --no-e2e --query-skew 0"
run poll_a X=1 -- $B
cp variants/librvcr_park.so $PKG/librvcr.so
( timeout 150 python -m pytest tests/test_gpu_coarse_f8.py -x -q ) > gpurun_out/r02n_pytest_f8.log 2>&1; echo "park f8 tests rc=$?"
run park_a X=1 -- $B
cp /tmp/librvcr_default.so $PKG/librvcr.so
run poll_b X=1 -- $B
cp variants/librvcr_park.so $PKG/librvcr.so
run park_b X=1 -- --steps 5 --no-e2e --query-skew 0
cp /tmp/librvcr_default.so $PKG/librvcr.so
