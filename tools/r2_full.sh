#!/bin/bash
# Round-2 GPU pass: parity tests, the bench line (FAST + STRICT + shapes), launch list, ncu --set full of the AMP kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
tail -3 gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
head -c 1500 gpurun_out/r2_bench_n1.json
