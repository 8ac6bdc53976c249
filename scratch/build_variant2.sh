#!/bin/bash
# usage: scratch/build_variant2.sh NAME file.cu -DFLAG ...   -> scratch/lib_NAME.so (only that file is recompiled)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift; shift
base=$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c kpgnn_b200/csrc/$src -o scratch/${base}_$name.o
objs=$(ls kpgnn_b200/build/*.o | grep -v "/$base.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/lib_$name.so $objs scratch/${base}_$name.o
