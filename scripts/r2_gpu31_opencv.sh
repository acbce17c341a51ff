#!/bin/bash
# round 2, GPU call 31: the CUDA path against the OpenCV-executed reference graphs (tests/golden/opencv_reference_graph_vectors.npz)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_vs_opencv_reference_gpu.py -q -x -p no:cacheprovider -s > gpurun_out/r2z_vs_opencv.log 2>&1; echo "rc=$?"; grep -E "OpenCV|passed|failed|Error" gpurun_out/r2z_vs_opencv.log | tail -8 | cut -c1-220
