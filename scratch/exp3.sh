#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --roofline-only 8192 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_launch'], d['frac'])"; }
run KP_LEAN_THREADS=1024
run KPGNN_B200_LIB=scratch/lib_pipe.so KP_LEAN_THREADS=1024
run KPGNN_B200_LIB=scratch/lib_pipe.so KP_LEAN_THREADS=256
KPGNN_B200_LIB=scratch/lib_pipe.so KP_LEAN_THREADS=1024 python scratch/dbg_lean.py 2>&1 | tail -12 | cut -c1-110
