#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 120 python scratch/small_batch.py 128 ${KS:-8} > gpurun_out/sb3_$tag.txt 2>&1; grep -E "full|gather|bwd" gpurun_out/sb3_$tag.txt; }
KS="1 2 4 8" run balanced FOO=1
KS="1 2 4 8" run unbalanced KP_LEAN_BALANCED=0
SECONDS=0
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_call3.json 2> gpurun_out/bench_call3.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_call3.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['backward']['frac'], d['roofline_batch128']['ms_per_launch'], d['roofline_batch128']['backward']['ms'])"
KP_LEAN_BALANCED=0 timeout 200 python bench.py --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('unbalanced', d['ms_per_step'], d['value'], d['e2e']['value'])"
