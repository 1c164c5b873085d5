#!/bin/bash
mkdir -p gpurun_out
echo "== regular"; timeout 90 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 1
echo "== no bar_ow (racy, timing only)"; SPARC_B200_LIB=build/lib_nobarow.so timeout 90 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 1
bash tools/r2_small.sh
P="python tools/profile_amp.py --T 8 --launches 3 --batch 296 --mode f64"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/r2_amp2_f64 $P > gpurun_out/r2_ncu_f64.log 2>&1
tail -2 gpurun_out/r2_ncu_f64.log
