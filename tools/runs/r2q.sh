#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_tests.log
tail -12 gpurun_out/r2q_tests.log
timeout 600 python tools/bench_configs.py auto cfg1,cfg2 > gpurun_out/r2q_configs.jsonl 2>&1
cat gpurun_out/r2q_configs.jsonl
