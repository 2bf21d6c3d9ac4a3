#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_interop.py -m gpu -q ) > gpurun_out/pytest_interop.log 2>&1
tail -30 gpurun_out/pytest_interop.log | cut -c1-300
( time timeout 900 python tools/build_probe.py --train ) > gpurun_out/build_probe.log 2>&1
tail -6 gpurun_out/build_probe.log | cut -c1-2400
