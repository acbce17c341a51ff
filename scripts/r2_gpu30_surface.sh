#!/bin/bash
# round 2, GPU call 30: the rebuilt library (new surface.cu: any-n SAME pooling, crop+concat, cross-entropy) through the whole operator suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2z_ops_surface.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2z_ops_surface.log | cut -c1-250
