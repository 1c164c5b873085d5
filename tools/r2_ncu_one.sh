#!/bin/bash
# one ncu --set full capture of the FAST pair kernel of the current build: tools/r2_ncu_one.sh TAG
P="python tools/profile_amp.py --T 8 --launches 3 --batch 296"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/r2_amp2_$1 $P > gpurun_out/r2_ncu_$1.log 2>&1
tail -2 gpurun_out/r2_ncu_$1.log
