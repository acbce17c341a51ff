#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r1b}
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 2300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
echo "== ncu launch list"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s ${2:-3600} -c ${3:-3400} --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; python scripts/summarize_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_summary.md 2>&1; head -45 gpurun_out/launches_${TAG}_summary.md
