#!/bin/bash
# final check of the round: full GPU suite, smoke, default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r5f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5f_tests.log
tail -3 gpurun_out/r5f_tests.log
python __graft_entry__.py smoke > gpurun_out/r5f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python bench.py > gpurun_out/r5f_bench.json 2> gpurun_out/r5f_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r5f_bench.json')); r = d['roofline']
print(d['ms_per_step'], r['kernel_ms_per_step'], r['frac'], r['traffic'], d['parity_check']['max_normwise_err'], d['clocks'])
PY
