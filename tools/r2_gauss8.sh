#!/bin/bash
N=8; A="--L 2048 --M 32 --rows 4096 --B 128 --T 20"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/gaussian_sharded.py $A --p2p --check --json gpurun_out/r2_gaussian_sharded_n${N}_p2p.json 2>&1 | grep -E "world|check" | tail -3
