#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python tools/kloop_hint_ab.py > gpurun_out/r5e_kloop_hint.log 2>&1
echo "rc=$?"; cat gpurun_out/r5e_kloop_hint.log | tail -8
