#!/bin/bash
# compute-sanitizer over every kernel family on small cases (golden clip, C1-sized synthetic), tensor-core paths forced:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# memcheck (out-of-bounds / misaligned / leaks of device memory), racecheck (shared-memory hazards between the
# warp-specialised roles), synccheck (barrier misuse), initcheck (reads of uninitialised device memory).  Summaries go
# to gpurun_out/sanitize_<tool>.log; copy them to profiles/ with the findings explained.
set -u
mkdir -p gpurun_out
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck initcheck}; do
  ( time timeout 420 compute-sanitizer --tool $tool --print-limit 40 --error-exitcode 99 \
      python tools/sanitize_case.py ) > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|case ok" gpurun_out/sanitize_$tool.log | tail -4
done
