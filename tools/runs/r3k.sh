#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_irfft.py > gpurun_out/r3k_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irfft_fast -c 1 -o gpurun_out/prof_r3k_irfft -f python tools/prof_irfft.py > gpurun_out/r3k_ncu.log 2>&1
tail -2 gpurun_out/r3k_ncu.log
