#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_stage3.py -m gpu -x -q -k "fused or golden" > gpurun_out/r3f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3f_tests.log
tail -4 gpurun_out/r3f_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-stage12 --no-other-configs > gpurun_out/r3f_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py --steps 20 --no-cpu --no-other-configs > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err
echo "bench rc=$?"
python - <<'PY'
import json, csv
d = json.load(open('gpurun_out/r3f_bench.json'))
print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['gpu_launches'], d.get('stage12_graph'))
rows=list(csv.reader(open('gpurun_out/r3f_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(r[ki][:60], float(r[vi].replace(',',''))/1e3) for r in rows[hi+2:] if len(r)>vi]
for s in seq[-12:]: print("%-62s %9.1f us"%s)
PY
