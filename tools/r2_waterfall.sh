#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/waterfall_vs_reference.py --n 4736 --flow soft --amp-mode fast --bp-mode fast --out gpurun_out/r2_waterfall_soft_fast.json 2>&1 | tail -12
timeout 500 python tools/waterfall_vs_reference.py --n 2368 --flow soft --amp-mode f64 --bp-mode strict --out gpurun_out/r2_waterfall_soft_f64.json 2>&1 | tail -12
