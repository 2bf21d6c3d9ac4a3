#!/bin/bash
# Round 2, GPU call J: coarse epilogue with a register queue and early hand-back of the accumulator stage; A/B of compile-time
# variants (tools/build_variants.sh) against the previous kernel on the same box.
set -u
mkdir -p gpurun_out
PKG=codename-rvc-fork-3_b200
cp $PKG/librvcr.so /tmp/librvcr_default.so
( time timeout 150 python -m pytest tests/test_gpu_coarse_tc.py tests/test_gpu_coarse_f8.py -x -q ) > gpurun_out/r02j_pytest_coarse.log 2>&1
rc=$?; echo "coarse rc=$rc" >> gpurun_out/r02j_pytest_coarse.log; tail -4 gpurun_out/r02j_pytest_coarse.log
if [ $rc -ne 0 ]; then echo "coarse tests failed: stopping"; tail -40 gpurun_out/r02j_pytest_coarse.log; exit 1; fi
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 python bench.py "$@" ) 2>gpurun_out/r02j_bench_$name.err | grep '^{' > gpurun_out/r02j_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02j_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()}, "clk", j.get("clocks", {}).get("sm_mhz"), j.get("config", {}).get("parity_sample", "")[:40], j.get("counters"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02j_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
B="--steps 3 --no-cpu-baseline --no-e2e --query-skew 0"
for v in s16q4 old s32q2 s16q2; do
  cp variants/librvcr_$v.so $PKG/librvcr.so
  run ${v}_e4m3 X=1 -- $B
  run ${v}_bf16 X=1 -- $B --coarse-f8-auto 0
done
cp /tmp/librvcr_default.so $PKG/librvcr.so
# tensor-pipe activity of the new default, e4m3
timeout 400 ncu --set full --clock-control none -k regex:coarse_tc_kernel -c 1 --csv --page raw --log-file gpurun_out/r02j_ncu_coarse.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --query-skew 0 > gpurun_out/r02j_ncu.log 2>&1
python - <<'PY'
import csv
try:
    rows = list(csv.reader(open("gpurun_out/r02j_ncu_coarse.csv")))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[h]
    for r in rows[h + 2:]:
        d = dict(zip(hdr, r))
        print(d["Kernel Name"][:50], {k: d[k] for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum") if k in d})
except Exception as e:
    print("ncu parse failed", e)
PY
