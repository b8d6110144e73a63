#!/bin/bash
# GPU session r2t (2 GPUs): NCCL / P2P / two-device tests, strong-scaling bench line at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_tests.log
tail -8 gpurun_out/r2t_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
echo "bench rc=$?"
tail -4 gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2t_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['scaling'])
print(json.dumps(d['scaling_variants'])[:2500])
print(json.dumps(d['e2e'])[:400])
PY
