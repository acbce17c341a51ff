#!/bin/bash
# round 2, GPU call 25: what bounds the fused tail forward kernel (generic vs register-tiled)? one ncu --set full launch of each
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for t5 in 0 1; do
PNP_TAIL5=$t5 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ps_mirror_conv -c 1 -o gpurun_out/r2x_tail$t5 python scripts/ncu_step.py --config 1 --steps 1 > gpurun_out/r2x_ncu_tail$t5.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2x_ncu_tail$t5.log
done
ls -la gpurun_out | grep r2x_tail
