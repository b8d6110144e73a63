#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for i in 1 2; do
for v in dbg il; do
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_$v.so timeout 200 python tools/fused_knobs.py 1024 3 "full $v:" "no stores $v:fused_debug=1" > gpurun_out/r5a_knobs_${v}_$i.log 2>&1
tail -2 gpurun_out/r5a_knobs_${v}_$i.log
done
done
