#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft.py -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_tests.log
tail -12 gpurun_out/r2m_tests.log
timeout 300 python tools/fft_bench.py > gpurun_out/r2m_fft_bench.log 2>&1
cat gpurun_out/r2m_fft_bench.log
