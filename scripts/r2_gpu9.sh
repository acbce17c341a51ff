#!/bin/bash
# round 2, GPU call 9: the complete GPU suite as the driver runs it + every bench config on the current build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu (whole suite)"; timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r2i_gpu_suite.log 2>&1; tail -4 gpurun_out/r2i_gpu_suite.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2i_gpu_suite.log | head
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for c in 4 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_c$c.json 2> gpurun_out/r2i_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2i_bench_c$c.json'));print('config $c', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e %.1f' % d['e2e']['value'], 'roof %.3f' % d['roofline']['frac'], d.get('n_D_20',{}).get('value'))"
done
grep -E "^\[(kern|conv)\]" gpurun_out/r2i_bench_c1.err | head
