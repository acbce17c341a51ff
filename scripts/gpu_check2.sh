#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 1500"
echo "== ops all"; timeout 1500 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/ops_all.log 2>&1; tail -4 gpurun_out/ops_all.log; grep -E "^FAILED" gpurun_out/ops_all.log
echo "== models all"; timeout 2400 $PY tests/test_models_gpu.py -m gpu > gpurun_out/models_all.log 2>&1; tail -4 gpurun_out/models_all.log; grep -E "^FAILED|worst|grad " gpurun_out/models_all.log | head -40
echo "== ncu launch list"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 3300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; python scripts/summarize_launches.py gpurun_out/launches_r1.csv > gpurun_out/launches_r1_summary.md 2>&1; head -40 gpurun_out/launches_r1_summary.md
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -c 2500 gpurun_out/bench_r1a.json
