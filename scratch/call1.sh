#!/bin/bash
# small-batch regime investigation (one gpurun call): A/B of launch geometry knobs + ncu source-level capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/sb_gpu.txt
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 120 python scratch/small_batch.py 128 ${KS:-8} > gpurun_out/sb_$tag.txt 2>&1; tail -6 gpurun_out/sb_$tag.txt; }
KS="1 2 4 8" run default FOO=1
run fwd_t128 KP_LEAN_THREADS=128
run fwd_t512 KP_LEAN_THREADS=512
run fwd_t1024 KP_LEAN_THREADS=1024
run b1_t256 KP_LEAN_B1_THREADS=256
run b1_t512 KP_LEAN_B1_THREADS=512
run pipe KPGNN_B200_LIB=scratch/lib_pipe.so
run pf1 KP_LEAN_PF_MODE=1
timeout 240 ncu --set full --clock-control none --cache-control none --import-source on \
  -k regex:"agg_fwd_lean|agg_bwd_dst_lean" -c 8 -f -o gpurun_out/sb128 python scratch/small_batch.py 128 8 --once > gpurun_out/sb_ncu.log 2>&1
tail -3 gpurun_out/sb_ncu.log
ls -la gpurun_out
