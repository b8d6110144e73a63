#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3e_bench_n2.json 2> gpurun_out/r3e_bench_n2.err
echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/r3e_bench_n2.json; wc -l gpurun_out/r3e_bench_n2.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r3e_bench_n2.json'))
print(d['value'], d['ms_per_step'], d.get('watchdog'), list(d.keys()))
print(d.get('stage12_graph'))
PY
