#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fused or cfg3 or shapes or many or host" > gpurun_out/r2v_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_tests.log
tail -5 gpurun_out/r2v_tests.log
timeout 600 python tools/fused_ab.py 1024 5 > gpurun_out/r2v_fused_ab.log 2>&1
cat gpurun_out/r2v_fused_ab.log
