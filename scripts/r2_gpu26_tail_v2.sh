#!/bin/bash
# round 2, GPU call 26: register-tiled 5x5 tail kernels with the nested-loop loader (v2) -- parity, then A/B against the generic kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== tail parity (PNP_TAIL5 default = 3)"
timeout 300 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "tail" -s > gpurun_out/r2y_tail.log 2>&1; echo "rc=$?"; grep -E "rel_err|passed|failed" gpurun_out/r2y_tail.log | tail -20 | cut -c1-200
for spec in "1 0" "1 1" "2 0" "2 2"; do set -- $spec; c=$1; t5=$2
  PNP_TAIL5=$t5 timeout 200 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-nd20 > gpurun_out/r2y_c${c}_tail$t5.json 2> gpurun_out/r2y_c${c}_tail$t5.err
  python -c "import json;d=json.load(open('gpurun_out/r2y_c${c}_tail$t5.json'));print('cfg$c tail5=$t5', '%.1f' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'])" || tail -5 gpurun_out/r2y_c${c}_tail$t5.err
  grep "simt:" gpurun_out/r2y_c${c}_tail$t5.err | grep -i "tail\|ps_" | head -4
done
