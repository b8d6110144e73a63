#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r3b_small.log 2>&1 <<'PY'
import sys, json, importlib.util, os
sys.argv = ['x']
import torch
sys.path.insert(0, os.getcwd())
import sfa_numpy_b200 as micarray
from oracle import sfa_oracle as O
spec = importlib.util.spec_from_file_location('bench', 'bench.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for prec in ("c64_fast", "c64"):
    for r in b.small_configs(torch, micarray, O, torch.device("cuda"), prec, 6561.6):
        sys.stderr.write(prec + " " + json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "note"}) + "\n")
PY
cat gpurun_out/r3b_small.log | tail -8
