#!/bin/bash
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "f64_mode or pair_kernel" -s 2>&1 | grep -E "f64 vs strict|passed|failed|Error|assert" | tail -20
for m in f64 strict fast; do
timeout 90 python tools/profile_amp.py --T 8 --launches 3 --batch 296 --mode $m 2>&1 | tail -n 1
done
timeout 90 python tools/profile_amp.py --T 64 --launches 2 --batch 2368 --mode f64 2>&1 | tail -n 1
timeout 90 python tools/profile_amp.py --T 64 --launches 2 --batch 2368 --mode strict 2>&1 | tail -n 1
