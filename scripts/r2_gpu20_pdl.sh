#!/bin/bash
# round 2, GPU call 20: programmatic dependent launch on every kernel of the library -- parity (eager and CUDA-graph replay), A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ops (PDL on)"
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2t_ops.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2t_ops.log | cut -c1-300
echo "== models / graph replay (PDL on)"
timeout 1200 python -m pytest tests/test_models_gpu.py tests/test_parity_configs_gpu.py -q -x -p no:cacheprovider -k "graph or config4 or config1 or config2_segmenter_train_step" > gpurun_out/r2t_models.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2t_models.log | cut -c1-300
for rep in 1 2; do
for pd in 0 1; do for pm in 0 1; do
  PNP_PDL=$pd PNP_TC_PAIR=$pm timeout 400 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_c4_pdl${pd}_pair${pm}_$rep.json 2> gpurun_out/r2t_c4_pdl${pd}_pair${pm}_$rep.err
  python -c "import json;d=json.load(open('gpurun_out/r2t_c4_pdl${pd}_pair${pm}_$rep.json'));print('cfg4 rep $rep pdl=$pd pair=$pm', '%.1f' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], 'roof %.3f' % d['roofline']['frac'])" || tail -5 gpurun_out/r2t_c4_pdl${pd}_pair${pm}_$rep.err
done; done; done
for pd in 0 1; do for c in 1 2 3; do
  PNP_PDL=$pd timeout 400 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_c${c}_pdl$pd.json 2> gpurun_out/r2t_c${c}_pdl$pd.err
  python -c "import json;d=json.load(open('gpurun_out/r2t_c${c}_pdl$pd.json'));print('cfg$c pdl=$pd', '%.1f' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'])" || tail -5 gpurun_out/r2t_c${c}_pdl$pd.err
done; done
