#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_irfft.py > gpurun_out/r2o_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irfft_fast -c 1 -o gpurun_out/prof_r2o_irfft -f python tools/prof_irfft.py > gpurun_out/r2o_ncu.log 2>&1
tail -3 gpurun_out/r2o_ncu.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_tests.log
tail -4 gpurun_out/r2o_tests.log
