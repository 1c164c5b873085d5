#!/bin/bash
# A/B timing of the AMP kernels at the headline shape: pair kernel vs one-codeword-per-CTA kernel
for p in 1 0; do
  echo "== SB_AMP_PAIR=$p"
  SB_AMP_PAIR=$p timeout 600 python tools/profile_amp.py --T 8 --launches 4 --batch 296 2>&1 | tail -2
  SB_AMP_PAIR=$p timeout 600 python tools/profile_amp.py --T 64 --launches 3 --batch 1184 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_ab.log
