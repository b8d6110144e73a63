#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 600 python tools/fused_attrib2.py 1024 > gpurun_out/r2j_attrib.log 2>&1
cat gpurun_out/r2j_attrib.log
