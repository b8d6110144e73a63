#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" _fft256x2 _fft256x3; do
  echo "== variant [$v]" >> gpurun_out/r3i_fft.log
  SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200$v.so timeout 300 python -m pytest tests/test_gpu_fft.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/r3i_fft.log
  SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200$v.so timeout 300 python tools/fft_bench.py 2>&1 | grep -v generic >> gpurun_out/r3i_fft.log
done
cat gpurun_out/r3i_fft.log
