#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in lb8 lb16 lb8 lb16; do
lib=$PWD/sfa_numpy_b200/lib/libsfa_b200_$v.so
echo "== $v"
SFA_B200_LIB=$lib timeout 100 python tools/fft_bench.py 2>&1 | grep "radix-16" | grep "irfft" >> gpurun_out/r5l_fft_$v.log
tail -5 gpurun_out/r5l_fft_$v.log
done
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_lb16.so timeout 200 python -m pytest tests/test_gpu_fft.py -m gpu -x -q > gpurun_out/r5l_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r5l_tests.log
