#!/bin/bash
# final pass of the round: smoke, GPU tests, bench, ncu of the FAST pair kernel
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
tail -3 gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"us/cwit",d["roofline"]["us_per_codeword_iteration"])
for k in ("f64","strict"):
    r=d.get(k,{}); print(k, r.get("value"), r.get("e2e"), r.get("mean_amp_iterations_per_decode"), r.get("roofline",{}).get("us_per_codeword_iteration"), r.get("roofline",{}).get("frac"))
print(d.get("speedup_vs_strict")); print(d.get("roofline_bp",{}).get("frac"), d.get("cpu_baseline",{}).get("value"))
for s in d.get("shapes",[]): print(s.get("shape"), s.get("codewords_per_s"), s.get("us_per_codeword_iteration"), s.get("frac"))
P
# (ncu capture: tools/r2_ncu_one.sh final)
