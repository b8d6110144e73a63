#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 200 python tools/cfg2_probe.py > gpurun_out/r4u_cfg2.log 2>&1
cat gpurun_out/r4u_cfg2.log
