#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/p2p_probe.py > gpurun_out/r2l_p2p_probe.log 2>&1
grep "rank" gpurun_out/r2l_p2p_probe.log
timeout 600 $TR --master-port 29522 tools/gather_probe.py 8 > gpurun_out/r2l_gather_probe.log 2>&1
grep "ms" gpurun_out/r2l_gather_probe.log | tail -12
tail -3 gpurun_out/r2l_gather_probe.log
