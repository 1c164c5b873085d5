#!/bin/bash
# Round-2 ncu captures of the current build: pair kernel (FAST), STRICT kernel, bench launch list.
mkdir -p gpurun_out
P="python tools/profile_amp.py --T 8 --launches 3 --batch 296"
timeout 600 $P > gpurun_out/r2_plain_fast.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/r2_amp2 $P > gpurun_out/r2_ncu_fast.log 2>&1
tail -2 gpurun_out/r2_ncu_fast.log
timeout 600 $P --mode strict > gpurun_out/r2_plain_strict.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_kernel -s 2 -c 1 -o gpurun_out/r2_amp_strict $P --mode strict > gpurun_out/r2_ncu_strict.log 2>&1
tail -2 gpurun_out/r2_ncu_strict.log
cat gpurun_out/r2_plain_fast.log gpurun_out/r2_plain_strict.log
B="python bench.py --steps 1 --warmup 1 --batch 2368 --streams 1 --no-cpu --no-strict --no-shapes"
timeout 600 $B > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log
