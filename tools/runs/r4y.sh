#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
n=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n \
    bench.py --gpus $n --steps 30 --warmup 3 > gpurun_out/r4y_bench_n$n.json 2> gpurun_out/r4y_bench_n$n.err
echo "bench n=$n rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r4y_bench_n$n.json'))
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['scaling'], d.get('watchdog'), d['parity_check']['max_normwise_err'])
    v = d['scaling_variants']
    for c in ('cfg3', 'cfg5'):
        print(c, {k: round(v[c][k]['ms'], 2) for k in v[c] if isinstance(v[c][k], dict) and 'ms' in v[c][k]})
    print('e2e', d['e2e']['value'], d['e2e'].get('frac_of_pcie'), d['e2e'].get('e2e_power_map',{}).get('value'))
except Exception as e:
    print('no line:', e)
PY
