#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python scripts/debug_gpu.py > gpurun_out/debug.log 2>&1; tail -60 gpurun_out/debug.log
PY="python -m pytest -p no:cacheprovider -q -rA --timeout 900"
echo "== models simt"; timeout 1500 $PY tests/test_models_gpu.py -m gpu -k "simt" > gpurun_out/models_simt.log 2>&1; grep -E "grad |worst|FAILED|passed|failed" gpurun_out/models_simt.log | head -80
