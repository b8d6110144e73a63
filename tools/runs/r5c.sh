#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_pwd.py 148 c64_fast > gpurun_out/r5c_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pwd_fused -c 1 -o gpurun_out/prof_r5c_fused -f python tools/prof_pwd.py 148 c64_fast > gpurun_out/r5c_ncu1.log 2>&1
tail -2 gpurun_out/r5c_ncu1.log; tail -1 gpurun_out/r5c_plain.log
