#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --roofline-only 8192 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_launch'], d['frac'])"; }
run KP_LEAN_PF_MODE=0 KP_LEAN_THREADS=256
run KP_LEAN_PF_MODE=0 KP_LEAN_THREADS=512
run KP_LEAN_PF_MODE=0 KP_LEAN_THREADS=1024
run KP_LEAN_PF_MODE=1 KP_LEAN_THREADS=1024
