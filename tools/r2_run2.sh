#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_kernel or fast or c3 or amp" -s > gpurun_out/r2_pair_test.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2_pair_test.log
grep -E "pair vs single|passed|failed" gpurun_out/r2_pair_test.log | tail -20
timeout 600 python tools/profile_amp.py --T 8 --launches 3 --batch 296 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/amp2_v1 python tools/profile_amp.py --T 8 --launches 3 --batch 296 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
