#!/bin/bash
# GPU session r2w (8 GPUs): strong-scaling bench line at N = 8 (and N = 4 on the same box)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2w_topo.txt 2>&1
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
    bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2w_bench_n$n.json 2> gpurun_out/r2w_bench_n$n.err
echo "bench n=$n rc=$?"
tail -3 gpurun_out/r2w_bench_n$n.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2w_bench_n$n.json'))
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['scaling'], d['parity_check'])
    print(json.dumps(d['scaling_variants'])[:2600])
    print(json.dumps(d['e2e'])[:900])
except Exception as e:
    print('no line:', e)
PY
done
