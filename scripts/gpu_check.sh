#!/bin/bash
# run on the GPU box through gpurun: staged so that a trapping tcgen05 kernel cannot take the SIMT results down with it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 900"
echo "== ops simt"; timeout 900 $PY tests/test_ops_gpu.py -m gpu -k "simt or not (auto or tc3 or tc1 or tensor_core or fused or dropout)" > gpurun_out/ops_simt.log 2>&1; tail -3 gpurun_out/ops_simt.log
echo "== ops tc unit"; timeout 900 $PY tests/test_ops_gpu.py -m gpu -k "tensor_core" > gpurun_out/ops_tc.log 2>&1; tail -3 gpurun_out/ops_tc.log
echo "== ops auto"; timeout 900 $PY tests/test_ops_gpu.py -m gpu -k "(auto or fused or dropout) and not tensor_core" > gpurun_out/ops_auto.log 2>&1; tail -3 gpurun_out/ops_auto.log
echo "== models simt"; timeout 1500 $PY tests/test_models_gpu.py -m gpu -k "simt" > gpurun_out/models_simt.log 2>&1; tail -3 gpurun_out/models_simt.log
echo "== models auto"; timeout 1500 $PY tests/test_models_gpu.py -m gpu -k "auto" > gpurun_out/models_auto.log 2>&1; tail -3 gpurun_out/models_auto.log
echo "== bench simt"; timeout 900 python bench.py --steps 3 --warmup 1 --backend simt --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; tail -c 600 gpurun_out/bench_simt.json
echo "== bench auto"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; tail -c 1500 gpurun_out/bench_auto.json; tail -5 gpurun_out/bench_auto.err
