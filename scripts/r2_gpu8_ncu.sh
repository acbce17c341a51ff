#!/bin/bash
# round 2, GPU call 8: ncu evidence (isolated call, bounded): launch list of two eager steps of config 4, one forward of config 1,
# and a --set full capture of the tcgen05 convolution kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== launch list config 4 (2 steps)"
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_c4.csv python scripts/ncu_step.py --config 4 --steps 2 > gpurun_out/r2h_ncu_c4.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2h_ncu_c4.log
python scripts/summarize_launches.py gpurun_out/r2h_launches_c4.csv > gpurun_out/r2h_launches_c4_summary.md 2>&1; head -30 gpurun_out/r2h_launches_c4_summary.md
echo "== launch list config 1 (forward stack, tensor-pipe % per launch)"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_c1.csv python scripts/ncu_step.py --config 1 --steps 1 > gpurun_out/r2h_ncu_c1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2h_ncu_c1.log
echo "== full capture of conv_tc kernels (config 4, first 16 launches of one step)"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 16 -o gpurun_out/r2h_conv_tc python scripts/ncu_step.py --config 4 --steps 1 > gpurun_out/r2h_ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2h_ncu_full.log
ls -la gpurun_out/ | grep r2h
