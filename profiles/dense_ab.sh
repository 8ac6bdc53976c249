#!/bin/bash
# A/B of the dense block's GEMM phases inside the captured training step (Kineto durations of the kernels in one replay):
# tensor-core 3xTF32 (default) vs fp32-FMA tiles (KP_DENSE_MMA=0), for a few rows-per-CTA choices.
for rows in ${ROWS:-36 32 48}; do for mma in 1 0; do
  KP_DENSE_ROWS=$rows KP_DENSE_MMA=$mma python profiles/step_timeline.py 2>/dev/null > /tmp/tl.txt
  echo "rows=$rows mma=$mma $(head -1 /tmp/tl.txt | cut -c1-80) $(grep -E '^#.*dense_block_(fwd|bwd)_kernel' /tmp/tl.txt | tr -s ' ' | tr '\n' ' ')"
done; done
