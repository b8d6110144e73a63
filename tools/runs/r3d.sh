#!/bin/bash
# final check (2 GPUs): whole GPU suite, smoke, bench lines at N = 1 and N = 2, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3d_tests.log
tail -3 gpurun_out/r3d_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r3d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3d_smoke.log
timeout 1200 python bench.py --steps 20 > gpurun_out/r3d_bench_n1.json 2> gpurun_out/r3d_bench_n1.err
echo "bench n1 rc=$?"; cut -c1-200 gpurun_out/r3d_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3d_bench_n2.json 2> gpurun_out/r3d_bench_n2.err
echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/r3d_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3d_bench_ref_n2.json 2> gpurun_out/r3d_bench_ref_n2.err
echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/r3d_bench_ref_n2.json
