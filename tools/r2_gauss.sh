#!/bin/bash
# Column-sharded Gaussian mode on N GPUs: tools/r2_gauss.sh N   (writes gpurun_out/r2_gaussian_sharded_n{N}[_p2p].json)
N=$1
mkdir -p gpurun_out
A="--L 2048 --M 32 --rows 4096 --B 128 --T 20"
if [ "$N" = "1" ]; then
  timeout 600 python tools/gaussian_sharded.py $A --json gpurun_out/r2_gaussian_sharded_n1.json 2>&1 | tail -2
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gaussian" -s 2>&1 | grep -E "Gaussian L=512|passed|failed|Error" | tail -16
else
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  timeout 600 $T tools/gaussian_sharded.py $A --json gpurun_out/r2_gaussian_sharded_n${N}.json 2>&1 | grep "world" | tail -1
  timeout 600 $T tools/gaussian_sharded.py $A --p2p --check --json gpurun_out/r2_gaussian_sharded_n${N}_p2p.json 2>&1 | grep -E "world|check" | tail -2
fi
