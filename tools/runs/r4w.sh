#!/bin/bash
# GPU session r4w: full GPU suite, bench line (default flags), launch list of the bench, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r4w_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4w_tests.log
tail -4 gpurun_out/r4w_tests.log
timeout 1200 python bench.py > gpurun_out/r4w_bench.json 2> gpurun_out/r4w_bench.err
echo "bench rc=$?"
tail -3 gpurun_out/r4w_bench.err
cut -c1-400 gpurun_out/r4w_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4w_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-stage12 --no-other-configs > gpurun_out/r4w_ncu2.log 2>&1
echo "launch list rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4w_bench_ref.json 2>> gpurun_out/r4w_bench.err
cut -c1-300 gpurun_out/r4w_bench_ref.json
python __graft_entry__.py smoke > gpurun_out/r4w_smoke.log 2>&1; echo "smoke rc=$?"
