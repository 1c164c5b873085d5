#!/bin/bash
# bench.py at N GPUs (FAST record + shapes; --no-strict): tools/r2_bench_n.sh N
N=$1; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $T bench.py --gpus $N --no-strict > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_final.json").readline())
print("N",d["n_gpus"],"value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"])
for s in d.get("shapes",[]): print(s.get("shape"), s.get("codewords_per_s"), s.get("frac"))
P
