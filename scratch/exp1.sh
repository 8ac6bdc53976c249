#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --roofline-only 8192 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_launch'], d['frac'])"; }
run KP_LEAN_PF_MODE=0
run KP_LEAN_PF_MODE=1 KP_LEAN_PF_DIST=1
run KP_LEAN_PF_MODE=1 KP_LEAN_PF_DIST=2
run KP_LEAN_PF_MODE=1 KP_LEAN_PF_DIST=4
run KP_LEAN_PF_MODE=2 KP_LEAN_PF_DIST=1
run KP_LEAN_PF_MODE=2 KP_LEAN_PF_DIST=2
run KPGNN_B200_LIB=scratch/lib_minb5.so KP_LEAN_PF_MODE=1
run KPGNN_B200_LIB=scratch/lib_minb6.so KP_LEAN_PF_MODE=1
run KPGNN_B200_LIB=scratch/lib_minb6.so KP_LEAN_PF_MODE=2 KP_LEAN_PF_DIST=2
