#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q ) > gpurun_out/pytest_fullsize.log 2>&1
tail -25 gpurun_out/pytest_fullsize.log | cut -c1-250
( time timeout 900 python tools/build_probe.py ) > gpurun_out/build_probe.log 2>&1
tail -6 gpurun_out/build_probe.log | cut -c1-1800
