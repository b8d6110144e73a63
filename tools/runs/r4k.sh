#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V='default: nohint:fused_debug=512 planes_unhinted:fused_debug=1024 both:fused_debug=1536 r16:fused_ring=16,fused_D=8 r16_nohint:fused_ring=16,fused_D=8,fused_debug=512'
for i in 1 2; do
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_dbg.so timeout 300 python tools/fused_knobs.py 1024 5 $V > gpurun_out/r4k_knobs_$i.log 2>&1
cat gpurun_out/r4k_knobs_$i.log
done
