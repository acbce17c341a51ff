#!/bin/bash
# round 2, GPU call 4: fixed tests + hand-over / evaluation tests + per-kernel profile (torch.profiler, no ncu)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 600"
echo "== ops"; timeout 900 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/r2d_ops.log 2>&1; tail -2 gpurun_out/r2d_ops.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2d_ops.log | head -30
echo "== handover"; timeout 900 $PY tests/test_handover_gpu.py -m gpu -s > gpurun_out/r2d_handover.log 2>&1; tail -2 gpurun_out/r2d_handover.log; grep -E "^(FAILED|ERROR)|^  " gpurun_out/r2d_handover.log | head -30
echo "== trajectory"; timeout 900 $PY tests/test_trajectory_gpu.py -m gpu -s -k "held_out or free_running_is" > gpurun_out/r2d_trajectory.log 2>&1; tail -2 gpurun_out/r2d_trajectory.log; grep -E "^(FAILED|ERROR)|held-out|logits max|after " gpurun_out/r2d_trajectory.log | head -30
echo "== models (quick: cfg4 + graph)"; timeout 900 $PY tests/test_models_gpu.py -m gpu -k "config4 or graph" > gpurun_out/r2d_models.log 2>&1; tail -2 gpurun_out/r2d_models.log
echo "== bench config 4 + profile"
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err
python -c "import json;d=json.load(open('gpurun_out/r2d_bench_c4.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('n_D_20',{}).get('value'))"
grep -E "^\[(kern|conv)\]" gpurun_out/r2d_bench_c4.err | head -12; grep -E "^\[prof\]" gpurun_out/r2d_bench_c4.err | head -45
echo "== bench config 5 (default / BK256=64)"
for v in 32 64; do
  PNP_TC_BK256=$v timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2d_bench_c5_bk$v.json 2> gpurun_out/r2d_bench_c5_bk$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2d_bench_c5_bk$v.json'));print('bk256=$v', {k:d[k] for k in ('value','ms_per_step')})"
done
for c in 2 3; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_c$c.json 2> gpurun_out/r2d_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2d_bench_c$c.json'));print('config $c', {k:d[k] for k in ('value','ms_per_step')})"
done
