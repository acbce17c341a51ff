#!/bin/bash
# round 2, GPU call 27: the NIfTI test protocol on the device (new test + the refactored per-volume evaluation), the generic-tail
# sub-process test, smoke, and the bench lines of configs 4 and 1 on the final tree
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_handover_gpu.py tests/test_modes_gpu.py -q -x -p no:cacheprovider -s -k "test_eval or per_volume or tail_parity" > gpurun_out/r2z_eval.log 2>&1; echo "rc=$?"; grep -E "agrees|per-class|passed|failed|Error|error" gpurun_out/r2z_eval.log | tail -20 | cut -c1-220
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 4 1; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2z_bench_c$c.json 2> gpurun_out/r2z_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2z_bench_c$c.json'));print('config $c', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e %.1f' % d['e2e']['value'], 'roof %.3f' % d['roofline']['frac'], d.get('n_D_20',{}).get('value'))" || tail -5 gpurun_out/r2z_bench_c$c.err
done
