#!/bin/bash
# round 2, GPU call 5 (2 GPUs): data-parallel parity test, dp_check in the bench line, bucketed vs single all-reduce
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "== test_dp_gpu"; timeout 600 python -m pytest -p no:cacheprovider -q -rA --timeout 500 tests/test_dp_gpu.py -m gpu -s > gpurun_out/r2e_dp_test.log 2>&1; tail -3 gpurun_out/r2e_dp_test.log; grep -E "max rel err" gpurun_out/r2e_dp_test.log
for nb in 4 0; do
  echo "== bench 2 GPUs, PNP_DP_BUCKETS=$nb"
  PNP_DP_BUCKETS=$nb NCCL_DEBUG=INFO timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-nd20 > gpurun_out/r2e_bench_2gpu_b$nb.json 2> gpurun_out/r2e_bench_2gpu_b$nb.err
  echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2e_bench_2gpu_b$nb.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d.get('dp_check'))"
  grep -E "NCCL INFO.*(nranks|Init COMPLETE|NVLS)" gpurun_out/r2e_bench_2gpu_b$nb.err | head -4
done
echo "== bench 1 GPU (same box) for the ratio"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench_1gpu.err
python -c "import json;d=json.load(open('gpurun_out/r2e_bench_1gpu.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')})"
