#!/bin/bash
# Round 2, GPU call D: coarse epilogue v3 with the top-8 floor + parked waits; the new feature tests.  Risky kernels first,
# every step under a short timeout of its own; a timeout in the first step ends the call.
set -u
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_coarse_tc.py -x -q ) > gpurun_out/r02d_pytest_coarse.log 2>&1
rc=$?; echo "coarse rc=$rc" >> gpurun_out/r02d_pytest_coarse.log; tail -6 gpurun_out/r02d_pytest_coarse.log
if [ $rc -eq 124 ]; then echo "coarse tests timed out: stopping"; exit 1; fi
( timeout 150 python -m pytest tests/test_gpu_coarse_f8.py -q -s ) > gpurun_out/r02d_pytest_f8.log 2>&1
echo "f8 rc=$?" >> gpurun_out/r02d_pytest_f8.log; grep -E "e4m3|passed|failed|rc=" gpurun_out/r02d_pytest_f8.log | tail -8
run() {  # name, bench args
  local name=$1; shift
  ( timeout 240 python bench.py "$@" ) 2>gpurun_out/r02d_bench_$name.err | grep '^{' > gpurun_out/r02d_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02d_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()},
          j.get("coarse_stats"), "parity", j["config"].get("parity_sample"), "clk", j.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02d_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
run c5_bf16 --steps 3 --no-cpu-baseline --no-e2e --query-skew 0
run c5_e4m3 --steps 3 --coarse-mode 4 --no-cpu-baseline --no-e2e --query-skew 0
( time timeout 900 python -m pytest tests/test_gpu_nprobe.py tests/test_gpu_interop.py tests/test_gpu_parity.py tests/test_gpu_final_proj.py tests/test_gpu_minibatch_kmeans.py tests/test_gpu_dropin.py tests/test_gpu_index_cache.py -q -s --timeout 300 ) > gpurun_out/r02d_pytest_new.log 2>&1
echo "rc=$?" >> gpurun_out/r02d_pytest_new.log; grep -E "passed|failed|rc=|^FAILED|^ERROR|inertia" gpurun_out/r02d_pytest_new.log | tail -25
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel' -c 1 -o gpurun_out/r02d_coarse_e4m3 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 --coarse-mode 4 ) > gpurun_out/r02d_ncu_e4m3.log 2>&1
ls -la gpurun_out | tail -4
