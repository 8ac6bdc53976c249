#!/bin/bash
mkdir -p gpurun_out
b() { env KP_LEAN_BALANCED=$1 timeout 100 python bench.py --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mask $1', d['ms_per_step'], d['value'], d['e2e']['value'])"; }
for r in 1 2 3; do for m in 0 3 2 1; do b $m; done; done | tee gpurun_out/balanced_ab.txt
