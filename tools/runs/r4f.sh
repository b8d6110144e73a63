#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_pipeline.py -m gpu -x -q -k "fused or cfg3 or padding or pipeline or fuzz or power or structure" > gpurun_out/r4f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4f_tests.log
tail -4 gpurun_out/r4f_tests.log
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 300 python tools/pad_ab.py 1024 3 > gpurun_out/r4f_pad_ab_dbg.log 2>&1
cat gpurun_out/r4f_pad_ab_dbg.log
