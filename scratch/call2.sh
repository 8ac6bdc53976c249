#!/bin/bash
# verification of the 64-entry window / batched staging: small-batch A/B, full GPU suite, bench line
mkdir -p gpurun_out
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 120 python scratch/small_batch.py 128 ${KS:-8} > gpurun_out/sb2_$tag.txt 2>&1; tail -${TL:-5} gpurun_out/sb2_$tag.txt; }
KS="1 2 4 8" TL=21 run default FOO=1
run fwd_t1024 KP_LEAN_THREADS=1024
run fwd_t512 KP_LEAN_THREADS=512
SECONDS=0
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -5 gpurun_out/pytest_gpu.log
SECONDS=0
timeout 200 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$? after ${SECONDS}s"; cat gpurun_out/bench_r1_final.json
