#!/bin/bash
# Round-2 final artefact pass (one GPU): full GPU test suite, smoke, the bench line, then -- only after the bench exited 0
# without ncu -- the ncu launch list of the same command (eager launches, see profile_r2.sh) and the step timeline.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=line 2>&1 | grep "AssertionError\|passed\|failed\|FAILED\|Error" | cut -c1-250 | tail -8 > gpurun_out/pytest_gpu_r2.log; cat gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 2>gpurun_out/bench_r2_final.err | tail -1 > gpurun_out/bench_r2_final.json || exit 1
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['backward']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --eager --steps 2 --warmup 3 --no-workloads --no-roofline --no-cpu-baseline > gpurun_out/ncu_launches_r2.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r2.csv
python profiles/step_timeline.py > gpurun_out/step_timeline_final.txt 2>/dev/null; head -1 gpurun_out/step_timeline_final.txt
