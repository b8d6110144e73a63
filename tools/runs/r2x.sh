#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in "" _sleep64 _sleep20k; do
  echo "== variant [$v] rep $rep" >> gpurun_out/r2x_sleep.log
  SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200$v.so timeout 300 python tools/prof_pwd.py 1024 c64_fast >> gpurun_out/r2x_sleep.log 2>&1
done
done
cat gpurun_out/r2x_sleep.log
