#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python tools/fused_knobs.py 148 2 "default:" "pair:fused_pair=1" > gpurun_out/r4o_small.log 2>&1
echo "small rc=$?"; cat gpurun_out/r4o_small.log | tail -5
timeout 200 python tools/fused_knobs.py 1024 5 "default:" "pair:fused_pair=1" > gpurun_out/r4o_knobs.log 2>&1
echo "knobs rc=$?"; tail -5 gpurun_out/r4o_knobs.log
nvidia-smi --query-gpu=name,memory.used --format=csv
