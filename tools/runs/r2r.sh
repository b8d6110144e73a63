#!/bin/bash
# GPU session r2r: full GPU suite, bench line (default flags), ncu capture of the fused kernel, launch list of the bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_tests.log
tail -4 gpurun_out/r2r_tests.log
timeout 1200 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"
tail -3 gpurun_out/r2r_bench.err
cut -c1-600 gpurun_out/r2r_bench.json
timeout 300 python tools/prof_pwd.py 148 c64_fast > gpurun_out/r2r_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pwd_fused -c 1 -o gpurun_out/prof_r2r_fused -f python tools/prof_pwd.py 148 c64_fast > gpurun_out/r2r_ncu1.log 2>&1
tail -2 gpurun_out/r2r_ncu1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-stage12 --no-other-configs > gpurun_out/r2r_ncu2.log 2>&1
echo "launch list rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_ref.json 2>> gpurun_out/r2r_bench.err
cut -c1-300 gpurun_out/r2r_bench_ref.json
