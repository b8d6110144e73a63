#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 100 python tools/fft_bench.py 2>&1 | grep "radix-16" > gpurun_out/r5m_fft.log
cat gpurun_out/r5m_fft.log
timeout 300 python -m pytest tests/test_gpu_fft.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r5m_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r5m_tests.log
