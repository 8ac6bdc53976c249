#!/bin/bash
# usage: scratch/build_variant.sh NAME -DKP_LEAN_MINB=5 ...   -> scratch/lib_NAME.so (only agg_fast_fwd.cu is recompiled)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c kpgnn_b200/csrc/agg_fast_fwd.cu -o scratch/fwd_$name.o
objs=$(ls kpgnn_b200/build/*.o | grep -v agg_fast_fwd.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/lib_$name.so $objs scratch/fwd_$name.o
cuobjdump -res-usage scratch/fwd_$name.o | grep -A1 "agg_fwd_lean_kernelILi32ELi1ELb1ELi1ELb0" | grep REG
