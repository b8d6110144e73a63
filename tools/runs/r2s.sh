#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/fused_ab.py 1024 5 > gpurun_out/r2s_fused_ab.log 2>&1
cat gpurun_out/r2s_fused_ab.log
timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_fullsize.py tests/test_gpu_extras.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_tests.log
tail -5 gpurun_out/r2s_tests.log
timeout 600 python - > gpurun_out/r2s_c128.log 2>&1 <<'PY'
import sys, json, importlib.util, os
sys.argv = ['x']
import torch
sys.path.insert(0, os.getcwd())
import sfa_numpy_b200 as micarray
from sfa_numpy_b200 import _lib
from oracle import sfa_oracle as O
spec = importlib.util.spec_from_file_location('bench', 'bench.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
lib = _lib.load()
geo = b.make_geometry(O, b.CFG)
for nd in (0.0, 1.0):
    lib.sfa_set_option(b"no_dmma", nd)
    rows = b.c128_rows(torch, micarray, O, torch.device("cuda"), b.CFG, geo)
    for r in rows:
        sys.stderr.write("no_dmma=%d %s\n" % (nd, json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})))
PY
cat gpurun_out/r2s_c128.log | tail -6
