#!/bin/bash
# Build an experiment variant of the library: tools/ab_build.sh TAG [extra nvcc flags...]
# Recompiles the M=512 instantiation (amp_inst_2.cu) and amp.cu with the flags and links build/lib_TAG.so from the
# other objects of the regular build.  Run a tool against it with SPARC_B200_LIB=build/lib_TAG.so.
set -e
cd "$(dirname "$0")/../sparc_ldpc_b200/csrc"
TAG=$1; shift
mkdir -p ../../build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
nvcc $F "$@" -c amp_inst_2.cu -o ../../build/inst2_$TAG.o 2>&1 | grep -A2 "amp_kernelILi9ELb1ELb1ELi1" | grep -E "spill|registers" || true
nvcc $F "$@" -c amp.cu -o ../../build/amp_$TAG.o > /dev/null 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_$TAG.so ../../build/amp_$TAG.o amp_inst_0.o amp_inst_1.o \
  ../../build/inst2_$TAG.o amp_inst_3.o bp.o handoff.o dense.o api.o -lcudart
echo built build/lib_$TAG.so
