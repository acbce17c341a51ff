#!/bin/bash
# round 2, GPU call 12: momentum optimizer + per-volume evaluation tests, dominant-kernel ncu capture (demangled name)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 600"
echo "== new tests"; timeout 900 $PY tests/test_ops_gpu.py tests/test_handover_gpu.py -m gpu -s -k "momentum or per_volume or evaluation_path or optimizers" > gpurun_out/r2l_new.log 2>&1; tail -3 gpurun_out/r2l_new.log; grep -E "^(FAILED|ERROR)|per-class" gpurun_out/r2l_new.log | head
echo "== full capture of conv_tc_kernel<256,...> (config 4)"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<\(int\)256" -c 8 -o gpurun_out/r2l_conv_tc256 python scripts/ncu_step.py --config 4 --steps 1 > gpurun_out/r2l_ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2l_ncu_full.log; ls -la gpurun_out | grep r2l_conv
