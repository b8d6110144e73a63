#!/bin/bash
# last check of the round: full GPU suite + smoke on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r5n_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5n_tests.log
tail -3 gpurun_out/r5n_tests.log
python __graft_entry__.py smoke > gpurun_out/r5n_smoke.log 2>&1; echo "smoke rc=$?"
