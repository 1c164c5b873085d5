#!/bin/bash
# On the GPU box: pair kernel of every experiment library given, balanced (T=8) and realistic (T=64) decodes
for t in "$@"; do
  echo "== $t"
  SPARC_B200_LIB=build/lib_$t.so timeout 600 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 1
  SPARC_B200_LIB=build/lib_$t.so timeout 600 python tools/profile_amp.py --T 64 --launches 2 --batch 4736 2>&1 | tail -n 1
done 2>&1 | tee gpurun_out/r2_ab3.log
