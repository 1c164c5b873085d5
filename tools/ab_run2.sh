#!/bin/bash
# On the GPU box: time the pair kernel of every experiment library given (tags of tools/ab_build2.sh)
for t in "$@"; do
  echo "== $t"
  SPARC_B200_LIB=build/lib_$t.so timeout 600 python tools/profile_amp.py --T 8 --launches 4 --batch 296 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_ab2.log
