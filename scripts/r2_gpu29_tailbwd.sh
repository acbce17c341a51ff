#!/bin/bash
# round 2, GPU call 29: the register-tiled tail BACKWARD on the step that actually runs it (config 4's G update; config 2 trains the
# output filter and never takes the fused tail path, so r2y's config-2 A/B said nothing about it)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for t5 in 1 3; do
  PNP_TAIL5=$t5 timeout 200 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu-baseline --no-nd20 > gpurun_out/r2z_c4_tail$t5.json 2> gpurun_out/r2z_c4_tail$t5.err
  python -c "import json;d=json.load(open('gpurun_out/r2z_c4_tail$t5.json'));print('cfg4 tail5=$t5', '%.1f' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'])" || tail -5 gpurun_out/r2z_c4_tail$t5.err
  grep "simt:" gpurun_out/r2z_c4_tail$t5.err | grep -i "tail" | head -4
done
