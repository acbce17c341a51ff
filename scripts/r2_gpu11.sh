#!/bin/bash
# round 2, GPU call 11: fused tail backward + final state: ops, models, handover, benches; demangled ncu capture of the dominant kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 600"
echo "== ops"; timeout 900 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/r2k_ops.log 2>&1; tail -2 gpurun_out/r2k_ops.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2k_ops.log | head -30
echo "== models"; timeout 900 $PY tests/test_models_gpu.py -m gpu > gpurun_out/r2k_models.log 2>&1; tail -2 gpurun_out/r2k_models.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2k_models.log | head
echo "== parity B=8 graph"; timeout 900 $PY tests/test_parity_configs_gpu.py -m gpu -k "b8" > gpurun_out/r2k_parity.log 2>&1; tail -2 gpurun_out/r2k_parity.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/r2k_$tag.json'));print('%-20s' % '$tag', '%.3f ms  %.1f slices/s' % (d['ms_per_step'], d['value']))"
}
run default PNP_X=0
run no_tail_bwd PNP_FUSE_TAIL_BWD=0
run default_again PNP_X=0
grep -E "^\[simt\]" gpurun_out/r2k_default.err | head -8
echo "== full capture of conv_tc_kernel<256,...> (config 4)"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<256" -c 8 -o gpurun_out/r2k_conv_tc256 python scripts/ncu_step.py --config 4 --steps 1 > gpurun_out/r2k_ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2k_ncu_full.log
