#!/bin/bash
# round 2, GPU call 22 (2 GPUs): final tree under torchrun -- DP parity test, bench with dp_check (CTA-pair kernels next to NCCL),
# the launch-mode tests (PNP_PDL=1, PNP_TC_PAIR=7 in sub-processes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== launch-mode tests"; timeout 900 python -m pytest -p no:cacheprovider -q --timeout 900 tests/test_modes_gpu.py -m gpu > gpurun_out/r2v_modes.log 2>&1; tail -3 gpurun_out/r2v_modes.log
echo "== test_dp_gpu"; timeout 600 python -m pytest -p no:cacheprovider -q --timeout 500 tests/test_dp_gpu.py -m gpu -s > gpurun_out/r2v_dp_test.log 2>&1; tail -3 gpurun_out/r2v_dp_test.log; grep -E "max rel err" gpurun_out/r2v_dp_test.log | head -4
echo "== bench 2 GPUs"
NCCL_DEBUG=INFO timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-nd20 > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err
echo "rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2v_bench_2gpu.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d.get('dp_check'))"
grep -E "NCCL INFO.*(nranks|Init COMPLETE|NVLS)" gpurun_out/r2v_bench_2gpu.err | head -3
echo "== bench 1 GPU (same box) for the ratio"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-nd20 > gpurun_out/r2v_bench_1gpu.json 2> gpurun_out/r2v_bench_1gpu.err
python -c "import json;d=json.load(open('gpurun_out/r2v_bench_1gpu.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['traffic'])"
