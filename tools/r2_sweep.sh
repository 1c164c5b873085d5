#!/bin/bash
for cfg in "9472 2" "9472 3" "9472 4" "18944 4" "18944 6"; do set -- $cfg
timeout 200 python bench.py --batch $1 --streams $2 --steps 2 --no-cpu --no-strict --no-shapes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch $1 streams $2: value %.0f e2e %.0f us/cwit %.3f frac %.4f'%(d['value'],d['e2e']['value'],d['roofline']['us_per_codeword_iteration'],d['roofline']['frac']))"
done
