#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -x -q > gpurun_out/r5j_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r5j_tests.log
