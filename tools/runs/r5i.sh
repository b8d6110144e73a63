#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" _g128 _g256p4; do
lib=$PWD/sfa_numpy_b200/lib/libsfa_b200$v.so
echo "== $lib"
SFA_B200_LIB=$lib timeout 200 python tools/fft_bench.py 2>&1 | grep "radix-16" | grep -v "n=4096" > gpurun_out/r5i_fft$v.log
cat gpurun_out/r5i_fft$v.log
done
