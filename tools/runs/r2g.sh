#!/bin/bash
# GPU session r2g: full GPU test suite, fused-kernel L2-hint / ring variants, new bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_tests.log
tail -3 gpurun_out/r2g_tests.log
for rep in 1 2; do
for v in "" _g2 _st _g2st _ring3 _g2ring3; do
  echo "== variant [$v] rep $rep" >> gpurun_out/r2g_variants.log
  SFA_B200_LIB=$PWD/sfa_numpy_b200/lib/libsfa_b200$v.so timeout 300 python tools/prof_pwd.py 1024 c64_fast >> gpurun_out/r2g_variants.log 2>&1
done
done
cat gpurun_out/r2g_variants.log
timeout 900 python bench.py --steps 10 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r2g_bench.err
cut -c1-1500 gpurun_out/r2g_bench.json
