#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/fused_knobs.py 1024 5 "default:" "no_cluster:no_cluster=1" > gpurun_out/r4r_knobs.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r4r_knobs.log
