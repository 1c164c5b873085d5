#!/bin/bash
# first GPU run of the pair kernel: parity tests, then A/B timing against the one-codeword-per-CTA kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_kernel" -s > gpurun_out/r2_pair_test.log 2>&1
echo "pair test rc=$?" | tee -a gpurun_out/r2_pair_test.log
tail -15 gpurun_out/r2_pair_test.log
for p in 1 0; do
  echo "== SB_AMP_PAIR=$p"
  SB_AMP_PAIR=$p timeout 600 python tools/profile_amp.py --T 8 --launches 4 --batch 296 2>&1 | tail -3
  SB_AMP_PAIR=$p timeout 600 python tools/profile_amp.py --T 64 --launches 3 --batch 1184 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r2_ab1.log
