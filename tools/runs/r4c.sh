#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_pipeline.py -m gpu -x -q -k "fused or cfg3 or padding or host or pipeline or fuzz or power" > gpurun_out/r4c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4c_tests.log
tail -4 gpurun_out/r4c_tests.log
timeout 900 python bench.py --steps 20 --no-cpu --no-stage12 --no-other-configs > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r4c_bench.json'))
print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
print(d.get('power_map_only'))
PY
