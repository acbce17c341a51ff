#!/bin/bash
# round 2, GPU call 7: A/B of the round's switches and of the K-block width of the 128/64-column tiles (bench config 4, 10 steps)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-nd20 > gpurun_out/r2g_$tag.json 2> gpurun_out/r2g_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/r2g_$tag.json'));print('%-28s' % '$tag', '%.3f ms  %.1f slices/s' % (d['ms_per_step'], d['value']))"
}
run default PNP_X=0
run bk128_32 PNP_TC_BK128=32
run bk64_32 PNP_TC_BK64=32
run bk128_64_both PNP_TC_BK128=32 PNP_TC_BK64=32
run no_fuse_epilogue PNP_FUSE_EPILOGUE=0
run no_fuse_tail PNP_FUSE_TAIL=0
run no_bn_direct PNP_BN_BWD_DIRECT=0
run no_planes_only PNP_PLANES_ONLY=0
run default_again PNP_X=0
timeout 300 python -m pytest -p no:cacheprovider -q --timeout 300 tests/test_ops_gpu.py -m gpu -k "fused_epilogue or tail_ps or conv_tensor_core" 2>&1 | tail -3
PNP_TC_BK128=32 PNP_TC_BK64=32 timeout 300 python -m pytest -p no:cacheprovider -q --timeout 300 tests/test_ops_gpu.py -m gpu -k "conv_tensor_core or conv_bn_relu or residual" 2>&1 | tail -3
