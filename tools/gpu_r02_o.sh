#!/bin/bash
# Round 2, GPU call O: the whole -m gpu suite and smoke() on the final build (the suite inside evidence call 3 stopped at a flaky
# assertion after 36 tests).
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r02o_pytest_gpu_full.log 2>&1; echo "full gpu suite rc=$?"; tail -6 gpurun_out/r02o_pytest_gpu_full.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02o_smoke.log
( timeout 300 python bench.py --steps 5 --no-e2e --query-skew 0 ) 2>/dev/null | grep '^{' > gpurun_out/r02o_bench.json
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02o_bench.json"))
print(round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j["phase_ms"].items()}, j["config"].get("parity_sample"))
PY
