#!/bin/bash
# Round 2, GPU call K: (1) is the streaming scan bound by bytes in flight?  RVCR_STC_NSTAGE sweep + a 2-tile-chunk build (7 stages);
# (2) centroid splits for the per-rank batch sizes of 4 / 8 GPUs (last-round quantisation of the m-blocks on 74 CTA pairs);
# (3) ncu of the main e4m3 coarse launch of the new epilogue.
set -u
mkdir -p gpurun_out
PKG=codename-rvc-fork-3_b200
cp $PKG/librvcr.so /tmp/librvcr_default.so
( time timeout 200 python -m pytest tests/test_gpu_coarse_tc.py tests/test_gpu_coarse_f8.py tests/test_gpu_scan_tc.py -x -q ) > gpurun_out/r02k_pytest.log 2>&1
rc=$?; echo "tests rc=$rc" >> gpurun_out/r02k_pytest.log; tail -4 gpurun_out/r02k_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed: stopping"; tail -40 gpurun_out/r02k_pytest.log; exit 1; fi
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 python bench.py "$@" ) 2>gpurun_out/r02k_bench_$name.err | grep '^{' > gpurun_out/r02k_bench_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/r02k_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j.get("phase_ms", {}).items()}, "clk", (j.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02k_bench_{sys.argv[1]}.err").read()[-1500:])
PY
}
B="--steps 3 --no-cpu-baseline --no-e2e --query-skew 0"
run default X=1 -- $B
for s in 3 4 5; do run nstage$s RVCR_STC_NSTAGE=$s -- $B; done
cp variants/librvcr_chunk2.so $PKG/librvcr.so
( timeout 150 python -m pytest tests/test_gpu_scan_tc.py -x -q ) > gpurun_out/r02k_pytest_chunk2.log 2>&1; echo "chunk2 scan tests rc=$?"
run chunk2 X=1 -- $B
cp /tmp/librvcr_default.so $PKG/librvcr.so
for q in 125000 250000; do
  for ns in 1 2 3 4; do run q${q}_ns$ns RVCR_TC_NSPLIT=$ns -- $B --queries $q; done
done
timeout 400 ncu --set full --clock-control none -k regex:coarse_tc_kernel -s 2 -c 1 --csv --page raw --log-file gpurun_out/r02k_ncu_coarse.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --query-skew 0 > gpurun_out/r02k_ncu.log 2>&1
python - <<'PY'
import csv
try:
    rows = list(csv.reader(open("gpurun_out/r02k_ncu_coarse.csv")))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[h]
    for r in rows[h + 2:]:
        d = dict(zip(hdr, r))
        print(d["Kernel Name"][:50], {k: d[k] for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum") if k in d})
except Exception as e:
    print("ncu parse failed", e)
PY
