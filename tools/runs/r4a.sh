#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/r4a_smi.log 2>&1
timeout 300 ./tools/store_probe5 > gpurun_out/r4a_probe5.log 2>&1
cat gpurun_out/r4a_probe5.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4a_tests.log
tail -4 gpurun_out/r4a_tests.log
