#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in p42 p24; do
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_$v.so timeout 200 python tools/fused_knobs.py 1024 4 "default:" "pair $v:fused_pair=1" > gpurun_out/r4p_knobs_$v.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r4p_knobs_$v.log
done
timeout 600 python -m pytest tests/test_gpu_stage3.py -m gpu -x -q -k "pair_mode" > gpurun_out/r4p_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r4p_tests.log
