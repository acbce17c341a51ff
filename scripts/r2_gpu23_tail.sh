#!/bin/bash
# round 2, GPU call 23: register-tiled 5x5 tail kernels (forward + backward) -- parity, A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== tail parity"
timeout 600 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "tail" -s > gpurun_out/r2w_tail.log 2>&1; echo "rc=$?"; grep -E "rel_err|passed|failed" gpurun_out/r2w_tail.log | tail -30 | cut -c1-250
for t5 in 0 3; do for c in 4 1; do
  PNP_TAIL5=$t5 timeout 400 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-nd20 > gpurun_out/r2w_c${c}_tail$t5.json 2> gpurun_out/r2w_c${c}_tail$t5.err
  python -c "import json;d=json.load(open('gpurun_out/r2w_c${c}_tail$t5.json'));print('cfg$c tail5=$t5', '%.1f' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'])" || tail -5 gpurun_out/r2w_c${c}_tail$t5.err
  grep "simt:" gpurun_out/r2w_c${c}_tail$t5.err | grep -i "tail\|ps_" | head -4
done; done
