#!/bin/bash
# A/B of programmatic dependent launch inside the captured step (Kineto, one replay; bench line for the step time).
for cfg in "0 0" "1 0" "1 1"; do set -- $cfg
  KP_DENSE_PDL=$1 KP_AGG_PDL=$2 python profiles/step_timeline.py 2>/tmp/tl.err > /tmp/tl.txt || tail -3 /tmp/tl.err
  echo "dense_pdl=$1 agg_pdl=$2 $(head -1 /tmp/tl.txt | cut -c1-80)"
  KP_DENSE_PDL=$1 KP_AGG_PDL=$2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads --no-roofline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   bench ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
