#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/fused_knobs.py 1024 5 "default:" "cluster4:fused_cluster=4" "prepass:fused_prepass=1" "cluster4 prepass:fused_cluster=4,fused_prepass=1" > gpurun_out/r4g_knobs.log 2>&1
cat gpurun_out/r4g_knobs.log
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 300 python tools/fused_knobs.py 1024 3 "full:" "full nohint:fused_debug=512" "stores only:fused_debug=26" "stores only nohint:fused_debug=538" "stores only, no conv:fused_debug=122" "stores only, no conv, nohint:fused_debug=634" "no MMA:fused_debug=8" "no MMA, no conv:fused_debug=104" "no operand loads:fused_debug=16" "cluster4 no MMA:fused_debug=8,fused_cluster=4" > gpurun_out/r4g_knobs_dbg.log 2>&1
cat gpurun_out/r4g_knobs_dbg.log
