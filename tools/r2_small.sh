#!/bin/bash
# small shapes with different CTA sizes (experiment library with the SB_AMP_THREADS knob)
for nt in 512 256 128 64; do
  echo "== SB_AMP_THREADS=$nt"
  SB_AMP_THREADS=$nt SPARC_B200_LIB=build/lib_exp.so timeout 120 python tools/bench_shapes.py --only C1,C4 --batch 9472 --reps 3 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(d['shape'], 'cw/s %.0f'%d['codewords_per_s'], 'us/cwit %.4f'%d['us_per_codeword_iteration'], 'frac %.4f'%d['frac'])
"
done
