#!/bin/bash
# round 2, GPU call 28: both training loops end to end on the device (event files, checkpoint/LR sequence); ncu --set full of the shipped
# register-tiled tail kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_handover_gpu.py -q -x -p no:cacheprovider -k "training_loops" > gpurun_out/r2z_loops.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2z_loops.log | cut -c1-250
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ps_mirror_conv5 -c 1 -o gpurun_out/r2z_tail5v2 python scripts/ncu_step.py --config 1 --steps 1 > gpurun_out/r2z_ncu_tail5v2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2z_ncu_tail5v2.log
