#!/bin/bash
# Round-2 profile pass (one GPU): the bench line, then -- only after it exited 0 without ncu -- the ncu launch list of the
# same command (short) and one --set full capture of the roofline launch (forward + backward aggregation kernels).
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r2.err | tail -1 > gpurun_out/bench_r2.json || exit 1
python -c "
import json; d=json.load(open('gpurun_out/bench_r2.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'])"
# --eager: the same kernels launched one by one instead of through the captured step graph (this ncu build aborts with
# "an error was reported by the driver" at the first graph launch of the round-2 step graph; under ncu every kernel is
# serialised anyway, so the list and the shares are the same)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --eager --steps 2 --warmup 3 --no-workloads --no-roofline --no-cpu-baseline > gpurun_out/ncu_launches_r2.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r2.csv
ncu --set full --clock-control none --import-source on -k regex:'agg_(fwd|bwd_dst)_lean_kernel' -c 4 -f -o gpurun_out/r2_agg \
  python bench.py --roofline-only 8192 > gpurun_out/ncu_agg_r2.log 2>&1
echo "set full rc=$?"; ls -la gpurun_out/r2_agg.ncu-rep
