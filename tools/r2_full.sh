#!/bin/bash
# Round-2 GPU pass: parity tests, the bench line (FAST + F64 + STRICT + shapes).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
tail -3 gpurun_out/r2_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_n1.err
python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"])
for k in ("f64","strict"):
    r=d.get(k,{}); print(k, r.get("value"), r.get("e2e"), r.get("mean_amp_iterations_per_decode"), r.get("roofline",{}).get("us_per_codeword_iteration"), r.get("roofline",{}).get("frac"))
print(d.get("speedup_vs_strict"))
for s in d.get("shapes",[]): print(json.dumps(s)[:330])
P
