#!/bin/bash
# round 2, GPU call 21: final state -- whole GPU suite, smoke, every bench config, reference arm, ncu evidence for the CTA-pair kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu (whole suite)"; timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r2u_gpu_suite.log 2>&1; tail -3 gpurun_out/r2u_gpu_suite.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2u_gpu_suite.log | head
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 4 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2u_bench_c$c.json 2> gpurun_out/r2u_bench_c$c.err
  python -c "import json;d=json.load(open('gpurun_out/r2u_bench_c$c.json'));print('config $c', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e %.1f' % d['e2e']['value'], 'roof %.3f' % d['roofline']['frac'], d['roofline']['kernel'][:32], 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d.get('n_D_20',{}).get('value'))" || tail -5 gpurun_out/r2u_bench_c$c.err
done
echo "== reference arm (config 4, same batch)"; timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err; python -c "import json;d=json.load(open('gpurun_out/r2u_bench_reference.json'));print({k:d[k] for k in ('impl','value','steps','warmup','ms_per_step')})"
echo "== ncu: launch list of two eager steps (config 4), --set full of 8 CTA-pair 128x256 launches"
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_c4.csv python scripts/ncu_step.py --config 4 --steps 2 > gpurun_out/r2u_ncu_c4.log 2>&1; echo "rc=$?"
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<\(int\)256" -c 8 -o gpurun_out/r2u_conv_tc256_pair python scripts/ncu_step.py --config 4 --steps 1 > gpurun_out/r2u_ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2u_ncu_full.log; ls -la gpurun_out | grep r2u_conv
