#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_stage3.py -m gpu -x -q -k "structure and c64_fast" > gpurun_out/r4z_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/r4z_memcheck.log
