#!/bin/bash
# GPU session r2k (2 GPUs): NCCL test of the sharded path, new grid test, strong-scaling bench line at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_stage3.py -m gpu -x -q -k "nccl or fused" > gpurun_out/r2k_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
tail -15 gpurun_out/r2k_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
echo "bench rc=$?"
tail -8 gpurun_out/r2k_bench_n2.err
cut -c1-3000 gpurun_out/r2k_bench_n2.json
