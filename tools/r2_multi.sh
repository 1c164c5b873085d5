#!/bin/bash
# Multi-GPU pass on N GPUs: column-sharded Gaussian mode (NCCL + peer memory), parity-mode drivers over NCCL, bench at N.
N=$1
mkdir -p gpurun_out
bash tools/r2_gauss.sh $N
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $T tools/mp_drivers_check.py --backend nccl --big 2> gpurun_out/r2_mp_n${N}.err | grep '^{' > gpurun_out/r2_mp_drivers_n${N}.jsonl; echo "mp rc=${PIPESTATUS[0]}"
cat gpurun_out/r2_mp_drivers_n${N}.jsonl | cut -c1-260
timeout 1200 $T bench.py --gpus $N --steps 3 --warmup 3 $BENCH_EXTRA > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_n${N}.json"))
print("N",d["n_gpus"],"value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"strict",d.get("strict",{}).get("value"))
for s in d.get("shapes",[]): print(json.dumps(s)[:300])
P
