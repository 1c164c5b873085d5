#!/bin/bash
# ncu --set full capture of one pair-kernel launch of an experiment library: tools/r2_ncu.sh TAG
t=$1
SPARC_B200_LIB=build/lib_$t.so timeout 600 python tools/profile_amp.py --T 8 --launches 3 --batch 296 > gpurun_out/plain_$t.log 2>&1 &&
SPARC_B200_LIB=build/lib_$t.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/amp2_$t python tools/profile_amp.py --T 8 --launches 3 --batch 296 > gpurun_out/ncu_$t.log 2>&1
tail -2 gpurun_out/ncu_$t.log
