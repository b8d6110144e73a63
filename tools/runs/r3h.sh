#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --no-other-configs --no-stage12 --no-cpu > gpurun_out/r3h_bench_n1.json 2> gpurun_out/r3h_bench_n1.err
echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu > gpurun_out/r3h_bench_n2.json 2> gpurun_out/r3h_bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r3h_bench_n1.json', 'gpurun_out/r3h_bench_n2.json'):
    d = json.load(open(f)); r = d['roofline']
    print(d['n_gpus'], d['value'], d['ms_per_step'], r['kernel_ms_per_step'], r['frac'], r['frac_step'], d['parity_check']['max_normwise_err'], d.get('watchdog'))
PY
