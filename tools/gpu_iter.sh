#!/bin/bash
# quick iteration call: new-kernel tests + phase timings at C3-like sizes
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_scan_tc.py -x -q -s ) > gpurun_out/pytest_scan_tc.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_scan_tc.log
( timeout 600 python tools/perf_probe.py --mode 2 --scan 5 ) > gpurun_out/probe_s5.log 2>&1
( timeout 600 python tools/perf_probe.py --mode 2 --scan 4 ) > gpurun_out/probe_s4.log 2>&1
tail -5 gpurun_out/pytest_scan_tc.log; tail -4 gpurun_out/probe_s5.log; tail -4 gpurun_out/probe_s4.log
