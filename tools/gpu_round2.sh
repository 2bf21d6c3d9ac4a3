#!/bin/bash
# Second call of a round: GPU parity tests (incl. the post-step and build-flow tests), smoke(), the bench line and
# the post-step probe.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
( time timeout 600 python tools/post_probe.py ) > gpurun_out/post_probe.log 2>&1
tail -3 gpurun_out/post_probe.log | cut -c1-1500
( time timeout 900 python bench.py ) > gpurun_out/bench_c5_n1_b.log 2>&1
grep '^{' gpurun_out/bench_c5_n1_b.log | cut -c1-400
