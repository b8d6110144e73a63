#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python tools/prof_irfft.py > gpurun_out/r5k_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:irfft_fast -c 1 -o gpurun_out/prof_r5k_irfft -f python tools/prof_irfft.py > gpurun_out/r5k_ncu.log 2>&1
tail -2 gpurun_out/r5k_ncu.log
