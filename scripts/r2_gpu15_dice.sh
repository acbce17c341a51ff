#!/bin/bash
# the held-out Dice gate three times (GPU training is run-to-run nondeterministic through split-K atomics)
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 900 python -m pytest tests/test_trajectory_gpu.py -m gpu -q -x -s -k dice_gate > gpurun_out/r2q_dice_$i.log 2>&1
  echo "run $i rc=$?"; grep -E "after|mean held|worst|passed|failed|logits max|evaluating" gpurun_out/r2q_dice_$i.log | cut -c1-230
done
