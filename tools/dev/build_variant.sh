#!/bin/bash
# Build a kernel-experiment library from the working tree: tools/dev/build_variant.sh NAME [-DFLAG ...]
# -> build_variants/lib_NAME.so (+ .ptxas with the register / spill report of every kernel)
name=$1; shift
cd "$(dirname "$0")/../../optical_networking_gym_b200/csrc" || exit 1
out=../../build_variants/lib_$name.so
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v "$@" \
  -o $out qrmsa_b200.cu tracegen.cpp 2> ../../build_variants/lib_$name.ptxas || { tail -20 ../../build_variants/lib_$name.ptxas; exit 1; }
grep -A1 "k_step_policyILi320ELi6ELi5ELi0ELi1ELi0E\|k_step_policyILi640ELi6ELi5ELi0ELi2ELi0E" ../../build_variants/lib_$name.ptxas | grep -v "^--" | sed 's/ptxas info    : //' | cut -c1-200
