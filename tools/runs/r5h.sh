#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fft.py tests/test_gpu_pipeline.py tests/test_gpu_extras.py -m gpu -x -q > gpurun_out/r5h_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r5h_tests.log
timeout 900 python bench.py --steps 10 --no-cpu --no-stage12 > gpurun_out/r5h_bench.json 2> gpurun_out/r5h_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r5h_bench.json'))
print(d['ms_per_step'], [(x['stage'][:12], round(x['ms'],3), round(x['frac_hbm'],3)) for x in d['fft']])
PY
