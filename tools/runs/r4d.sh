#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/pad_ab.py 1024 5 > gpurun_out/r4d_pad_ab.log 2>&1
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 300 python tools/pad_ab.py 1024 3 > gpurun_out/r4d_pad_ab_dbg.log 2>&1
cat gpurun_out/r4d_pad_ab.log gpurun_out/r4d_pad_ab_dbg.log
