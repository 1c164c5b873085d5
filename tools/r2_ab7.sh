#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_kernel or amp_c3 or fast or randomised or amp_trace or power_alloc" 2>&1 | tail -3
timeout 90 python tools/profile_amp.py --T 8 --launches 3 --batch 296 2>&1 | tail -n 1
timeout 90 python tools/profile_amp.py --T 64 --launches 2 --batch 4736 2>&1 | tail -n 1
timeout 120 python tools/bench_shapes.py --only C1,C4 --batch 9472 --reps 3 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(d['shape'], 'cw/s %.0f'%d['codewords_per_s'], 'us/cwit %.4f'%d['us_per_codeword_iteration'], 'frac %.4f'%d['frac'])
"
