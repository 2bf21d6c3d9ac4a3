#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -40 gpurun_out/pytest_gpu3.log
