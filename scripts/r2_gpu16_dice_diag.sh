#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python tests/diag_dice_gate.py --scale 0.25 --rounds 12 --decay 1.0 --check 4 > gpurun_out/r2r_diag_$i.log 2>&1; echo "$i rc=$?"; grep -E "^round|re-labelled|Error|error" gpurun_out/r2r_diag_$i.log | cut -c1-300
done
