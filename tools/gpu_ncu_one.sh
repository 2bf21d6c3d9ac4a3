#!/bin/bash
# one --set full capture of kernels matching $1 (regex) from a short C5 bench run; extra bench flags in $2..
set -u
mkdir -p gpurun_out
re=$1; shift
( time timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -c 1 -o gpurun_out/ncu_one -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" ) > gpurun_out/ncu_one.log 2>&1
tail -3 gpurun_out/ncu_one.log
