#!/bin/bash
# Build an experiment variant of the pair kernel: tools/ab_build2.sh TAG [extra nvcc flags...]
# Recompiles amp2.cu with the flags and links build/lib_TAG.so from the other objects of the regular build.
# Run a tool against it with SPARC_B200_LIB=build/lib_TAG.so.
set -e
cd "$(dirname "$0")/../sparc_ldpc_b200/csrc"
TAG=$1; shift
mkdir -p ../../build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
nvcc $F "$@" -c amp2.cu -o ../../build/amp2_$TAG.o 2>&1 | grep -A2 "amp2_kernel" | grep -E "spill|registers" || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_$TAG.so amp.o ../../build/amp2_$TAG.o amp_inst_0.o amp_inst_1.o \
  amp_inst_2.o amp_inst_3.o bp.o handoff.o dense.o api.o -lcudart
echo built build/lib_$TAG.so
