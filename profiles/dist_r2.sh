#!/bin/bash
# Round-2 two-GPU verification (run with `gpurun --gpus 2`): bounded legs, each under its own timeout, progress in
# gpurun_out/dist_rank*.log.  Legs: bench with the collective outside the step graph; the gradient / replica check both
# ways; bench with the collective captured in the graph.
mkdir -p gpurun_out; rm -f gpurun_out/dist_rank*.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
KP_NCCL_IN_GRAPH=0 timeout -k 5 240 $TR --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-workloads 2>gpurun_out/b_n2_out.err | tail -1 > gpurun_out/b_n2_out.json; echo "bench out-of-graph rc=$?"
KP_NCCL_IN_GRAPH=0 timeout -k 5 150 $TR --master-port 29513 tests/dist_check.py 2>gpurun_out/dc0.err | tail -2; echo "check out-of-graph rc=${PIPESTATUS[0]}"
KP_NCCL_IN_GRAPH=1 timeout -k 5 150 $TR --master-port 29514 tests/dist_check.py 2>gpurun_out/dc1.err | tail -2; echo "check in-graph rc=${PIPESTATUS[0]}"
KP_NCCL_IN_GRAPH=1 timeout -k 5 200 $TR --master-port 29515 bench.py --gpus 2 --steps 30 --warmup 5 --no-workloads 2>gpurun_out/b_n2_in.err | tail -1 > gpurun_out/b_n2_in.json; echo "bench in-graph rc=$?"
for f in gpurun_out/b_n2_out.json gpurun_out/b_n2_in.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])
except Exception as e: print('$f', 'no line', e)"; done
tail -n 30 gpurun_out/dist_rank0.log
