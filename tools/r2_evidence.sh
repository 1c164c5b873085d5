#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/cliff_point.py --n 4736 --k 256 --json gpurun_out/r2_cliff_point.json > gpurun_out/r2_cliff.log 2>&1; echo "cliff rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/r2_cliff_point.json"))
a=d["A_same_codewords_three_modes"]; print(a["ok"], a["comparison"])
for k,v in a["modes"].items(): print(k, v["block_failures_per_stage"], v["ber_per_stage[amp1,ldpc1,amp2,ldpc2,amp3]"][:2])
b=d["B_reference_stream_vs_oracle"]; print(b["ok"], {k:(v["identical_tuples"],v["different"],v["unexplained"]) for k,v in b["modes"].items()})
P
timeout 300 python tools/waterfall_vs_reference.py --n 4736 --flow soft --amp-mode fast --bp-mode fast --out gpurun_out/r2_waterfall_soft_fast.json 2>&1 | tail -5 | cut -c1-200
