#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py -m gpu -x -q -k "fused or power or plan" > gpurun_out/r2y_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2y_tests.log
tail -5 gpurun_out/r2y_tests.log
timeout 900 python bench.py --steps 10 --no-cpu --no-stage12 --no-other-configs > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2y_bench.json'))
print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d.get('power_map_only'), d['collective'])
PY
