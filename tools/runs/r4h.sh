#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 300 python tools/fused_knobs.py 1024 5 "default:" "D8 r24:fused_D=8,fused_ring=24" "D16 r64:fused_D=16,fused_ring=64" "D24 r48:fused_D=24,fused_ring=48" "D32 r96:fused_D=32,fused_ring=96" "plane stores unhinted:fused_debug=1024" "prepass:fused_prepass=1" > gpurun_out/r4h_knobs_dbg.log 2>&1
cat gpurun_out/r4h_knobs_dbg.log
timeout 300 python tools/prof_pwd.py 148 c64_fast > gpurun_out/r4h_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pwd_fused -c 1 -o gpurun_out/prof_r4h_fused -f python tools/prof_pwd.py 148 c64_fast > gpurun_out/r4h_ncu1.log 2>&1
tail -2 gpurun_out/r4h_ncu1.log
