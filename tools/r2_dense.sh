#!/bin/bash
for c in 32 64 128; do for cl in 2 4; do
echo "== SB_DENSE_CHUNK_KB=$c CLUSTER=$cl"
SB_DENSE_CHUNK_KB=$c SB_DENSE_CLUSTER=$cl SPARC_B200_LIB=build/lib_dexp.so timeout 120 python tools/profile_dense.py --reps 3 --check 2>&1 | grep -E "max err|rep 2"
done; done
