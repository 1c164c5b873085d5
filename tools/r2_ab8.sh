#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_kernel or amp_c3 or fast or randomised or amp_trace or power_alloc or c5 or flows" 2>&1 | tail -3
timeout 90 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 1
timeout 90 python tools/profile_amp.py --T 64 --launches 2 --batch 4736 2>&1 | tail -n 1
timeout 90 env SB_AMP_PAIR=0 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 1
