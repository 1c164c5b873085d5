#!/bin/bash
# extra randomised parity sweeps with seeds the GPU tests do not use (tests use 3 / 1 / 2 / 1); log -> profiles/r02_fuzz_extra_seeds.log
mkdir -p gpurun_out
LOG=gpurun_out/r2_fuzz_extra.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run python tools/fuzz_parity.py --cases 400 --seed 11 --budget-s 70
run python tools/fuzz_flows.py 7 150
run python tools/fuzz_bp.py 5
run python tools/fuzz_handoff.py 5
grep -E "^==|^rc=|fuzz|EXCEEDED|identical|cases" $LOG | cut -c1-400
