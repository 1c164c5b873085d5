#!/bin/bash
# Experiment build of the whole AMP part with the environment knobs compiled in (-DSB_EXPERIMENT): tools/ab_build3.sh TAG [flags]
# -> build/lib_TAG.so; run a tool against it with SPARC_B200_LIB=build/lib_TAG.so [SB_AMP_THREADS=256 ...]
set -e
cd "$(dirname "$0")/../sparc_ldpc_b200/csrc"
TAG=$1; shift
mkdir -p ../../build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DSB_EXPERIMENT"
for f in amp amp2 amp_inst_0 amp_inst_1 amp_inst_2 amp_inst_3; do nvcc $F "$@" -c $f.cu -o ../../build/${f}_$TAG.o & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_$TAG.so ../../build/amp_$TAG.o ../../build/amp2_$TAG.o \
  ../../build/amp_inst_0_$TAG.o ../../build/amp_inst_1_$TAG.o ../../build/amp_inst_2_$TAG.o ../../build/amp_inst_3_$TAG.o bp.o handoff.o dense.o api.o -lcudart
echo built build/lib_$TAG.so
