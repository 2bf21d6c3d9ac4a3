#!/bin/bash
# Round 2, evidence call (1 GPU): the judged artefacts of the shipped build.
#   bench line (default flags), reference arm, C4 line, ncu launch list of the bench command, ncu --set full of the dominant
#   kernels, compute-sanitizer over every kernel family.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_smi.txt 2>&1
( time timeout 400 python bench.py ) > gpurun_out/r02_bench_c5_n1.log 2>&1
grep '^{' gpurun_out/r02_bench_c5_n1.log > gpurun_out/r02_bench_c5_n1.json; tail -c 300 gpurun_out/r02_bench_c5_n1.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c5_n1_steps20.json
if [ -z "${SLIM:-}" ]; then   # SLIM=1: only the C5 lines and the ncu captures (a re-run after a small change)
( timeout 300 python bench.py --coarse-f8-auto 0 --steps 5 --no-cpu-baseline --query-skew 0 ) 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c5_n1_bf16.json
( timeout 300 python bench.py --impl reference --steps 2 ) 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c5_ref.json
( timeout 400 python bench.py --workload C4 --steps 3 ) 2>/dev/null | grep '^{' > gpurun_out/r02_bench_c4.json
for w in C1 C2 C3; do ( timeout 200 python bench.py --workload $w --steps 20 --query-skew 0 ) 2>/dev/null | grep '^{' > gpurun_out/r02_bench_${w}.json; done
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        j = json.load(open(f))
        print(f.split("/")[-1], round(j["ms_per_step"], 3), "ms", round(j["value"]), j["unit"], "e2e", j.get("e2e") and round(j["e2e"]["value"]),
              "roof", j.get("roofline") and (j["roofline"].get("dominant"), round(j["roofline"]["frac"], 3)), "parity", (j.get("config") or {}).get("parity_sample"))
    except Exception as e:
        print(f, "FAILED", e)
PY
( time timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:'rvcr::|DeviceRadixSort|DeviceSelect' -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 ) > gpurun_out/r02_ncu_list.log 2>&1
( time timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'coarse_tc_kernel|scan_tc_kernel|scan_recheck_kernel|tc_recheck_kernel' \
    -s 16 -c 4 -o gpurun_out/r02_full_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --query-skew 0 ) > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_full_c5.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_c5.csv 2>/dev/null
if [ -n "${SLIM:-}" ]; then ls -la gpurun_out | tail -6; exit 0; fi
bash tools/sanitize.sh
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/r02_pytest_gpu_full.log 2>&1; echo "full gpu suite rc=$?"; tail -5 gpurun_out/r02_pytest_gpu_full.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
ls -la gpurun_out | tail -6
