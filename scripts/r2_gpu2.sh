#!/bin/bash
# round 2, GPU call 2: full GPU suite on the new epilogue / fused-BN kernels, bench, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 1500"
echo "== ops"; timeout 1200 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/r2b_ops.log 2>&1; tail -4 gpurun_out/r2b_ops.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2b_ops.log | head -30
echo "== models"; timeout 1500 $PY tests/test_models_gpu.py -m gpu > gpurun_out/r2b_models.log 2>&1; tail -4 gpurun_out/r2b_models.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2b_models.log | head
echo "== parity configs"; timeout 1500 $PY tests/test_parity_configs_gpu.py -m gpu -s > gpurun_out/r2b_parity_configs.log 2>&1; tail -4 gpurun_out/r2b_parity_configs.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2b_parity_configs.log
echo "== trajectory"; timeout 1500 $PY tests/test_trajectory_gpu.py -m gpu -s > gpurun_out/r2b_trajectory.log 2>&1; tail -4 gpurun_out/r2b_trajectory.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2b_trajectory.log
for c in 4 1; do
  echo "== bench config $c"
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c$c.json 2> gpurun_out/r2b_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2b_bench_c$c.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('n_D_20',{}).get('value'))"
  grep -E "^\[(kern|conv)\]" gpurun_out/r2b_bench_c$c.err | head -12
done
echo "== launch list (2 eager steps, config 4)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-nd20 > gpurun_out/r2b_bench_under_ncu.json 2> gpurun_out/r2b_bench_under_ncu.err
python scripts/summarize_launches.py gpurun_out/r2b_launches.csv > gpurun_out/r2b_launches_summary.md 2>&1; head -40 gpurun_out/r2b_launches_summary.md
