#!/bin/bash
# Round-2 multi-GPU verification (run with `gpurun --gpus N`, N = $1 or 2): bounded legs, each under its own timeout,
# progress in gpurun_out/dist_rank*.log.  Legs: gradient / replica check with the peer-memory exchange; bench with the
# peer-memory exchange (default); bench with the process group's all-reduce between two graphs (KP_PEER_ALLREDUCE=0).
N=${1:-2}
mkdir -p gpurun_out; rm -f gpurun_out/dist_rank*.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 5 200 $TR --master-port 29513 tests/dist_check.py 2>gpurun_out/dc_peer.err | tail -2; echo "check peer rc=${PIPESTATUS[0]}"
timeout -k 5 200 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-workloads 2>gpurun_out/b_n${N}_peer.err | tail -1 > gpurun_out/b_n${N}_peer.json; echo "bench peer rc=$?"
KP_PEER_ALLREDUCE=0 timeout -k 5 200 $TR --master-port 29515 bench.py --gpus $N --steps 30 --warmup 5 --no-workloads 2>gpurun_out/b_n${N}_nccl.err | tail -1 > gpurun_out/b_n${N}_nccl.json; echo "bench nccl rc=$?"
for f in gpurun_out/b_n${N}_peer.json gpurun_out/b_n${N}_nccl.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config'].get('gradient_exchange'))
except Exception as e: print('$f', 'no line', e)"; done
tail -n 4 gpurun_out/dist_rank0.log; grep -i "warn.*peer\|unavailable" gpurun_out/*.err | head -5
