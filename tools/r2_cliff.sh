#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/cliff_point.py --n 4736 --k 256 --json gpurun_out/r2_cliff_point.json > gpurun_out/r2_cliff.log 2>&1; echo "cliff rc=$?"
tail -60 gpurun_out/r2_cliff.log | cut -c1-200
