#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" _f256 _f384; do
lib=$PWD/sfa_numpy_b200/lib/libsfa_b200$v.so
echo "== $lib"
SFA_B200_LIB=$lib timeout 200 python tools/fft_bench.py 2>&1 | grep "radix-16" > gpurun_out/r5g_fft$v.log
cat gpurun_out/r5g_fft$v.log
done
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_f256.so timeout 300 python -m pytest tests/test_gpu_fft.py -m gpu -x -q > gpurun_out/r5g_tests_f256.log 2>&1
echo "pytest f256 rc=$?"; tail -2 gpurun_out/r5g_tests_f256.log
