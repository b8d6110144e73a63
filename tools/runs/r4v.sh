#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fuzz.py tests/test_gpu_pipeline.py -m gpu -x -q -k "fused or structure or pipeline or fuzz" > gpurun_out/r4v_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r4v_tests.log; grep -n "^E " gpurun_out/r4v_tests.log | head -5
timeout 300 python tools/bench_configs.py auto cfg2 > gpurun_out/r4v_cfg.log 2>&1
cat gpurun_out/r4v_cfg.log
