#!/bin/bash
# Round 2, after tools/gpu_r02_first.sh has shown which variant wins: the judged artefacts for that variant.
#   BENCH_ARGS="--coarse-mode 4 --scan-impl 6" gpurun --timeout 1500 -- 'bash tools/gpu_r02_ncu.sh'
# bench line (not under a profiler), reference arm, ncu launch list of the same command, one --set full capture of the
# dominant kernels.  Copy the summaries into profiles/r02_*.
set -u
A=${BENCH_ARGS:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_smi.txt 2>&1
( time timeout 900 python bench.py $A ) > gpurun_out/r02_bench_c5_n1.log 2>&1
tail -c 400 gpurun_out/r02_bench_c5_n1.log
( time timeout 600 python bench.py --impl reference --steps 2 ) > gpurun_out/r02_bench_ref.log 2>&1
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:'rvcr::|DeviceRadixSort|DeviceSelect' -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $A ) > gpurun_out/r02_ncu_list.log 2>&1
( time timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'coarse_tc_kernel|scan_tc_kernel|scan_tc_fused_kernel|scan_recheck_kernel|tc_recheck_kernel' \
    -c 4 -o gpurun_out/r02_full_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e $A ) > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_full_c5.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_c5.csv 2>/dev/null
ls -la gpurun_out | tail -12
