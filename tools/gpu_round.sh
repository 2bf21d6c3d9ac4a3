#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench command and one
# --set full capture of the two dominant kernels.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_c5_n1.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 ) > gpurun_out/bench_ref.log 2>&1
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'rvcr::|DeviceRadixSort|DeviceSelect' -c 400 --csv \
    --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/ncu_list.log 2>&1
( time timeout 900 ncu --set full --clock-control none --import-source on -k regex:'coarse_tc_kernel|scan_tc_kernel|scan_recheck_kernel|tc_recheck_kernel' \
    -c 4 -o gpurun_out/r01_full_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
