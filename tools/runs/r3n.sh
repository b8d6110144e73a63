#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r3n_small.log 2>&1 <<'PY'
import sys, json, importlib.util, os
sys.argv = ['x']
import torch
sys.path.insert(0, os.getcwd())
import sfa_numpy_b200 as micarray
from sfa_numpy_b200 import _lib
from oracle import sfa_oracle as O
lib = _lib.load()
spec = importlib.util.spec_from_file_location('bench', 'bench.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for nc in (0.0, 1.0):
    lib.sfa_set_option(b"no_cluster", nc)
    for r in b.small_configs(torch, micarray, O, torch.device("cuda"), "c64_fast", 6561.6):
        sys.stderr.write("no_cluster=%d " % nc + json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("config", "ms", "frac_hbm")}) + "\n")
PY
cat gpurun_out/r3n_small.log | tail -6
