#!/bin/bash
# Build a variant of the library with extra -D flags: tools/build_alt.sh <name> [-DKNOB=value ...] -> build/alt/lib_<name>.so
# (select it with SPART_B200_LIB=build/alt/lib_<name>.so; tools/kbench.py prints per-kernel times and a checksum)
set -e
name=$1; shift
mkdir -p build/alt
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o build/alt/lib_$name.so spart-python_b200/csrc/spart_kernels.cu
