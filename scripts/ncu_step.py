"""Profiling target for ncu: builds one bench workload and runs eager steps, with the profiled range marked by
cudaProfilerStart/Stop (ncu --profile-from-start off).  No CUDA graphs, no NCCL, no nvidia-smi polling, no CPU baseline.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/ncu_step.py --config 4 --steps 2
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 16 -o gpurun_out/conv_tc \
        python scripts/ncu_step.py --config 4 --steps 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=None)
    a = ap.parse_args()
    a.batch = a.batch or bench.WORKLOADS[a.config][1]
    a.backend = "tc1" if a.config == 5 else "auto"
    a.keep_prob, a.graph = 0.75, False
    torch.cuda.set_device(0)
    w = bench.Workload(a, torch.device("cuda", 0), 0, 1)
    for i in range(a.warmup):
        w.step_resident(i, eager=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(a.steps):
        w.step_resident(i, eager=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ncu_step: config %d, %d profiled eager step(s) at B=%d" % (a.config, a.steps, a.batch))


if __name__ == "__main__":
    main()
