#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage3.py -m gpu -x -q -k "fused_kernel_shapes" > gpurun_out/r5b_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r5b_tests.log; grep -n "^E " gpurun_out/r5b_tests.log | head -5
