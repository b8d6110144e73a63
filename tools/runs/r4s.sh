#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py auto,steering,factored cfg1,cfg2 > gpurun_out/r4s_cfg.log 2>&1
cat gpurun_out/r4s_cfg.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r4s_launches.csv python tools/bench_configs.py factored cfg1,cfg2 > gpurun_out/r4s_ncu.log 2>&1
echo "ncu rc=$?"
