#!/bin/bash
# final multi-GPU check (8 GPUs): bench lines at N = 8 and N = 4, reference arm under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for n in 8 4; do
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n \
    bench.py --gpus $n --steps 30 --warmup 3 > gpurun_out/r3m_bench_n$n.json 2> gpurun_out/r3m_bench_n$n.err
echo "bench n=$n rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r3m_bench_n$n.json'))
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['scaling'], d.get('watchdog'), d['parity_check']['max_normwise_err'])
    v = d['scaling_variants']
    for c in ('cfg3', 'cfg5'):
        print(c, {k: round(v[c][k]['ms'], 2) for k in v[c] if isinstance(v[c][k], dict) and 'ms' in v[c][k]})
    print('e2e', d['e2e']['value'], d['e2e'].get('frac_of_pcie'))
except Exception as e:
    print('no line:', e)
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 \
    bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r3m_ref_n8.json 2> gpurun_out/r3m_ref_n8.err
echo "ref rc=$?"; cut -c1-160 gpurun_out/r3m_ref_n8.json
