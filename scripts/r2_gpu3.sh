#!/bin/bash
# round 2, GPU call 3: staged and bounded (call 2 lost its box): new-kernel op tests first, then the suites, then benches. No ncu here.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 600"
free -g | head -2; nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv
echo "== new-kernel ops"; timeout 600 $PY tests/test_ops_gpu.py -m gpu -k "fused_epilogue or tail_ps or dropout or fused_bn_stats or conv_bn_relu" > gpurun_out/r2c_ops_new.log 2>&1; tail -3 gpurun_out/r2c_ops_new.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2c_ops_new.log | head -20
echo "== ops all"; timeout 900 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/r2c_ops.log 2>&1; tail -3 gpurun_out/r2c_ops.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2c_ops.log | head -30
echo "== models"; timeout 900 $PY tests/test_models_gpu.py -m gpu > gpurun_out/r2c_models.log 2>&1; tail -3 gpurun_out/r2c_models.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2c_models.log | head
free -g | head -2
echo "== bench config 4"
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
python -c "import json;d=json.load(open('gpurun_out/r2c_bench_c4.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('n_D_20',{}).get('value'))"
grep -E "^\[(kern|conv)\]" gpurun_out/r2c_bench_c4.err | head -12
echo "== parity configs"; timeout 900 $PY tests/test_parity_configs_gpu.py -m gpu -s > gpurun_out/r2c_parity_configs.log 2>&1; tail -3 gpurun_out/r2c_parity_configs.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2c_parity_configs.log
free -g | head -2
echo "== trajectory"; timeout 900 $PY tests/test_trajectory_gpu.py -m gpu -s > gpurun_out/r2c_trajectory.log 2>&1; tail -3 gpurun_out/r2c_trajectory.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2c_trajectory.log
free -g | head -2
echo "== bench config 1"
timeout 600 python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c1.json 2> gpurun_out/r2c_bench_c1.err
python -c "import json;d=json.load(open('gpurun_out/r2c_bench_c1.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])"
echo "== input pipeline"; timeout 120 python scripts/bench_input_pipeline.py --seconds 2 --out gpurun_out/r2c_input_pipeline.json 2>&1 | tail -7
