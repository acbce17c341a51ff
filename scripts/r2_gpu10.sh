#!/bin/bash
# round 2, GPU call 10: validate the resident-weight variant of the narrow tcgen05 tiles (+ A/B), full capture of the dominant kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 600"
echo "== ops (BRES on)"; timeout 900 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/r2j_ops.log 2>&1; tail -2 gpurun_out/r2j_ops.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2j_ops.log | head -30
echo "== models cfg1/cfg4"; timeout 900 $PY tests/test_models_gpu.py -m gpu -k "config1 or config4" > gpurun_out/r2j_models.log 2>&1; tail -2 gpurun_out/r2j_models.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2j_models.log | head
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2j_$tag.json 2> gpurun_out/r2j_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/r2j_$tag.json'));print('%-20s' % '$tag', '%.3f ms  %.1f slices/s' % (d['ms_per_step'], d['value']))"
  grep -E "^\[kern\] conv_tc_kernel<(16|32)," gpurun_out/r2j_$tag.err | head -6
}
run bres_on PNP_TC_BRES=1
run bres_off PNP_TC_BRES=0
run bres_on_again PNP_TC_BRES=1
echo "== full capture of conv_tc_kernel<256,...> (config 4)"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:conv_tc_kernel<256" -c 8 -o gpurun_out/r2j_conv_tc256 python scripts/ncu_step.py --config 4 --steps 1 > gpurun_out/r2j_ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2j_ncu_full.log
