#!/bin/bash
# GPU session r2i: stream-ahead conversion in the fused kernel -- correctness, then timing against the pre-pass
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage3.py -m gpu -x -q -k "fused" > gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
tail -15 gpurun_out/r2i_tests.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "cfg3" >> gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
tail -5 gpurun_out/r2i_tests.log
for rep in 1 2; do
  echo "== stream-ahead rep $rep" >> gpurun_out/r2i_timing.log
  timeout 200 python tools/prof_pwd.py 1024 c64_fast >> gpurun_out/r2i_timing.log 2>&1
  echo "== prepass rep $rep" >> gpurun_out/r2i_timing.log
  SFA_FUSED_PREPASS=1 timeout 200 python tools/prof_pwd.py 1024 c64_fast >> gpurun_out/r2i_timing.log 2>&1
done
cat gpurun_out/r2i_timing.log
