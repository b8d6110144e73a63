#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_extras.py -m gpu -x -q > gpurun_out/r4m_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4m_tests.log
tail -3 gpurun_out/r4m_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r4m_bench_n2.json 2> gpurun_out/r4m_bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r4m_bench_n2.json')); r = d['roofline']
print(d['n_gpus'], d['value'], d['ms_per_step'], r['kernel_ms_per_step'], r['frac'], d['parity_check']['max_normwise_err'], d.get('watchdog'))
v = d['scaling_variants']
for c in ('cfg3', 'cfg5'):
    print(c, {k: round(v[c][k]['ms'], 2) for k in v[c] if isinstance(v[c][k], dict) and 'ms' in v[c][k]})
print('e2e', d['e2e']['value'], d['e2e'].get('frac_of_pcie'))
PY
