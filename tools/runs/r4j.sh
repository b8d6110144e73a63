#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_pipeline.py -m gpu -x -q -k "fused or cfg3 or padding or pipeline or fuzz or power or structure" > gpurun_out/r4j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4j_tests.log
tail -3 gpurun_out/r4j_tests.log
V='default: D8r32:fused_D=8,fused_ring=32 D16r40:fused_D=16,fused_ring=40 D16r32:fused_D=16,fused_ring=32'
for i in 1 2; do
timeout 300 python tools/fused_knobs.py 1024 5 $V > gpurun_out/r4j_knobs_$i.log 2>&1
cat gpurun_out/r4j_knobs_$i.log
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_early.so timeout 300 python tools/fused_knobs.py 1024 5 $V > gpurun_out/r4j_knobs_early_$i.log 2>&1
echo early; tail -5 gpurun_out/r4j_knobs_early_$i.log
done
