#!/bin/bash
# round 2, GPU call 1: new parity tests at the benchmarked configs + baseline bench lines of every --config (before kernel work)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 1500"
echo "== parity configs"; timeout 1500 $PY tests/test_parity_configs_gpu.py -m gpu -s > gpurun_out/r2a_parity_configs.log 2>&1; tail -5 gpurun_out/r2a_parity_configs.log
echo "== trajectory"; timeout 1500 $PY tests/test_trajectory_gpu.py -m gpu -s > gpurun_out/r2a_trajectory.log 2>&1; tail -5 gpurun_out/r2a_trajectory.log
for c in 4 1 2 3 5; do
  echo "== bench config $c"
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c$c.json 2> gpurun_out/r2a_bench_c$c.err
  tail -c 700 gpurun_out/r2a_bench_c$c.json; grep -E "^\[(kern|conv)\]" gpurun_out/r2a_bench_c$c.err | head -12
done
