#!/bin/bash
# round 2, GPU call 13 (8 GPUs): the scaling points the driver will run, with dp_check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
for n in $N 1; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
  else
    NCCL_DEBUG=INFO timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $n --steps 10 --warmup 3 --no-nd20 > gpurun_out/r2m_bench_n$n.json 2> gpurun_out/r2m_bench_n$n.err
  fi
  echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2m_bench_n$n.json'));print('N=$n', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e %.1f' % d['e2e']['value'], d.get('dp_check'))"
done
grep -E "NCCL INFO.*(nranks|NVLS multicast)" gpurun_out/r2m_bench_n$N.err | head -3
