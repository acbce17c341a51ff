#!/bin/bash
# round 2, GPU call 14: final state -- whole GPU suite, smoke, every bench config, step-level DRAM capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu (whole suite)"; timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r2n_gpu_suite.log 2>&1; tail -3 gpurun_out/r2n_gpu_suite.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2n_gpu_suite.log | head
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 4 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2n_bench_c$c.json 2> gpurun_out/r2n_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2n_bench_c$c.json'));print('config $c', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e %.1f' % d['e2e']['value'], 'roof %.3f' % d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d.get('n_D_20',{}).get('value'))"
done
echo "== reference arm (config 4, same batch)"; timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; python -c "import json;d=json.load(open('gpurun_out/r2n_bench_reference.json'));print({k:d[k] for k in ('impl','value','steps','warmup','ms_per_step')}, d['cpu_baseline']['sample'])"
echo "== step-level DRAM traffic (config 4, one eager step; config 1, one forward)"
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2n_dram_c4.csv python scripts/ncu_step.py --config 4 --steps 1 > gpurun_out/r2n_ncu_dram_c4.log 2>&1; echo "rc=$?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_c1.csv python scripts/ncu_step.py --config 1 --steps 1 > gpurun_out/r2n_ncu_c1.log 2>&1; echo "rc=$?"
