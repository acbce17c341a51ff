#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r1}
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 1500"
echo "== ops"; timeout 1500 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/ops_$TAG.log 2>&1; tail -2 gpurun_out/ops_$TAG.log; grep -E "^FAILED" gpurun_out/ops_$TAG.log
echo "== models"; timeout 2400 $PY tests/test_models_gpu.py -m gpu > gpurun_out/models_$TAG.log 2>&1; tail -2 gpurun_out/models_$TAG.log; grep -E "^FAILED|worst" gpurun_out/models_$TAG.log | head -30
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches'],d['roofline']['achieved'],d['roofline']['kernel_ms_per_step'],d['cpu_baseline'])"; grep "\[tc\]" gpurun_out/bench_$TAG.err | head -8
