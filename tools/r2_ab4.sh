#!/bin/bash
# A/B of the current library: pair-kernel parity test + lockstep (T=8) and realistic (T=64) timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_kernel or amp_c3 or fast_mode" -s 2>&1 | grep -E "pair vs single|passed|failed|Error" | tail -12
timeout 600 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 2
timeout 600 python tools/profile_amp.py --T 64 --launches 2 --batch 4736 2>&1 | tail -n 1
