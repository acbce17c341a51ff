#!/bin/bash
# round 2, GPU call 19: CTA-pair (cta_group::2) convolution kernel -- parity of all three pair tile shapes, A/B against the
# single-CTA kernels on configs 4 / 1 / 5; the re-designed Dice gate three times
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pair kernel parity (PNP_TC_PAIR=7: N 256 / 128 / 64)"
PNP_TC_PAIR=7 timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "tensor_core or cta_pair or fused_epilogue or residual or conv_bn" > gpurun_out/r2s_ops_pair7.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s_ops_pair7.log | cut -c1-300
echo "== default (PNP_TC_PAIR=1) ops + parity configs"
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "tensor_core or cta_pair or fused_epilogue" > gpurun_out/r2s_ops_default.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s_ops_default.log | cut -c1-300
for pm in 0 1 7; do
  PNP_TC_PAIR=$pm timeout 400 python bench.py --config 4 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2s_c4_pair$pm.json 2> gpurun_out/r2s_c4_pair$pm.err
  python -c "import json;d=json.load(open('gpurun_out/r2s_c4_pair$pm.json'));print('cfg4 pair=$pm', d['value'], d['ms_per_step'], d['roofline']['kernel'][:34], '%.3f' % d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])" || tail -5 gpurun_out/r2s_c4_pair$pm.err
done
for pm in 0 1; do
  for c in 1 5; do
  PNP_TC_PAIR=$pm timeout 400 python bench.py --config $c --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2s_c${c}_pair$pm.json 2> gpurun_out/r2s_c${c}_pair$pm.err
  python -c "import json;d=json.load(open('gpurun_out/r2s_c${c}_pair$pm.json'));print('cfg$c pair=$pm', d['value'], d['ms_per_step'], d['roofline']['kernel'][:34], '%.3f' % d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])" || tail -5 gpurun_out/r2s_c${c}_pair$pm.err
  done
done
echo "== Dice gate x3"
for i in 1 2 3; do
  timeout 900 python -m pytest tests/test_trajectory_gpu.py -m gpu -q -x -s -k dice_gate > gpurun_out/r2s_dice_$i.log 2>&1
  echo "run $i rc=$?"; grep -E "after|mean held|worst|passed|failed|evaluating" gpurun_out/r2s_dice_$i.log | cut -c1-230
done
