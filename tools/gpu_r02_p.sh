#!/bin/bash
# Round 2, GPU call P: coarse re-check with ballot-compacted candidates (four per pass).
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_coarse_tc.py tests/test_gpu_coarse_f8.py tests/test_gpu_parity.py tests/test_gpu_build_flow.py -x -q ) > gpurun_out/r02p_pytest.log 2>&1
rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/r02p_pytest.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/r02p_pytest.log; exit 1; fi
for v in e4m3 bf16; do
  extra=""; [ $v = bf16 ] && extra="--coarse-f8-auto 0"
  ( timeout 300 python bench.py --steps 5 --no-e2e --query-skew 0 $extra ) 2>gpurun_out/r02p_bench_$v.err | grep '^{' > gpurun_out/r02p_bench_$v.json
  python - $v <<'PY'
import json, sys
j = json.load(open(f"gpurun_out/r02p_bench_{sys.argv[1]}.json"))
print(sys.argv[1], round(j["ms_per_step"], 3), "ms/step", {k: round(v, 2) for k, v in j["phase_ms"].items()}, j["config"].get("parity_sample"), j.get("coarse_stats"))
PY
done
( timeout 200 python bench.py --workload C3 --steps 10 --no-e2e --query-skew 0 ) 2>/dev/null | grep '^{' | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('C3', round(j['ms_per_step'],3), {k: round(v, 2) for k, v in j['phase_ms'].items()}, j['config'].get('parity_sample'))"
