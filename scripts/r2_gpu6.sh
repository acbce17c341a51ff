#!/bin/bash
# round 2, GPU call 6: BN backward without the g round trip, planes-only hidden activations, fast tail loader
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 600"
echo "== ops"; timeout 900 $PY tests/test_ops_gpu.py -m gpu > gpurun_out/r2f_ops.log 2>&1; tail -2 gpurun_out/r2f_ops.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2f_ops.log | head -30
echo "== models"; timeout 900 $PY tests/test_models_gpu.py -m gpu > gpurun_out/r2f_models.log 2>&1; tail -2 gpurun_out/r2f_models.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2f_models.log | head
echo "== handover"; timeout 900 $PY tests/test_handover_gpu.py -m gpu -s > gpurun_out/r2f_handover.log 2>&1; tail -2 gpurun_out/r2f_handover.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2f_handover.log | head
echo "== parity (B=8 graph, B=16) + trajectory (adversarial, teacher-forced)"; timeout 900 $PY tests/test_parity_configs_gpu.py tests/test_trajectory_gpu.py -m gpu -s -k "b8 or b16 or adversarial_trajectory or teacher_forced or graph_replay_tracks" > gpurun_out/r2f_parity.log 2>&1; tail -2 gpurun_out/r2f_parity.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2f_parity.log | head
echo "== bench config 4 + profile"
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err
python -c "import json;d=json.load(open('gpurun_out/r2f_bench_c4.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('n_D_20',{}).get('value'))"
grep -E "^\[prof\]" gpurun_out/r2f_bench_c4.err | head -32
for c in 2 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2f_bench_c$c.json 2> gpurun_out/r2f_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2f_bench_c$c.json'));print('config $c', {k:d[k] for k in ('value','ms_per_step')})"
done
