#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V='default: D8r24:fused_D=8,fused_ring=24 D8r16:fused_D=8,fused_ring=16 D12r30:fused_D=12,fused_ring=30 D4r12:fused_D=4,fused_ring=12'
for i in 1 2; do
timeout 300 python tools/fused_knobs.py 1024 5 $V > gpurun_out/r4i_knobs_$i.log 2>&1
cat gpurun_out/r4i_knobs_$i.log
SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200_early.so timeout 300 python tools/fused_knobs.py 1024 5 $V > gpurun_out/r4i_knobs_early_$i.log 2>&1
echo early; cat gpurun_out/r4i_knobs_early_$i.log
done
