#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fft.py -m gpu -x -q -k "host or fft or irfft or stft or examples" > gpurun_out/r3l_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3l_tests.log
tail -4 gpurun_out/r3l_tests.log
timeout 900 python bench.py --steps 10 --no-cpu --no-stage12 --no-other-configs > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r3l_bench.json'))
print(d['ms_per_step'], d['e2e']['value'], d['e2e']['e2e_power_map'])
PY
