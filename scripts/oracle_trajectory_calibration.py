"""How far does an fp32 CPU reference drift from an fp64 one over a multi-step trajectory?

Runs the oracle (oracle/pnp_graphs.py) in fp32 and fp64 from identical initial variables and inputs for
N optimizer steps and prints per-step losses of both.  The trajectory parity tests
(tests/test_trajectory_gpu.py) hold the CUDA path to the fp32 oracle; this script documents the noise floor
an fp32 implementation of the same math has against exact arithmetic.

    python scripts/oracle_trajectory_calibration.py seg 12
    python scripts/oracle_trajectory_calibration.py adv 12

Each run merges its per-step losses (full precision) into tests/golden/oracle_trajectory_calibration.json, the fixture
tests/test_trajectory_gpu.py reads.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.pnp_graphs import (OracleSegmenter, OracleAdversarial, init_numpy_params, synthetic_images,  # noqa: E402
                               synthetic_labels)
from oracle.tf14_numpy import label_decomp  # noqa: E402


def seg(n, B=2):
    ws, bns = OracleSegmenter.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    x = synthetic_images(B, 1234)
    y = torch.from_numpy(label_decomp(5, synthetic_labels(B, 99)))
    o32 = OracleSegmenter(P, B)
    o64 = OracleSegmenter(P, B, dtype=torch.float64)
    rec = {"wce32": [], "wce64": [], "dice32": [], "dice64": []}
    for s in range(n):
        t0 = time.time()
        a = o32.train_step(x, y, 1.0)
        b = o64.train_step(x.double(), y.double(), 1.0)
        print("seg step %2d  wce %.7f / %.7f (rel %.2e)  dice %.7f / %.7f (rel %.2e)  %.1fs" % (
            s, a["wce"], b["wce"], abs(a["wce"] - b["wce"]) / abs(b["wce"]), a["dice"], b["dice"],
            abs(a["dice"] - b["dice"]) / abs(b["dice"]), time.time() - t0), flush=True)
        for k, v in (("wce32", a["wce"]), ("wce64", b["wce"]), ("dice32", a["dice"]), ("dice64", b["dice"])):
            rec[k].append(v)
    return rec


def adv(n, B=2):
    ws, bns = OracleAdversarial.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    for nme, s in ws:
        if "cls" in nme:
            P[nme] = np.clip(P[nme] * 0.5, -0.05, 0.05).astype(np.float32)
    kw = dict(lambda_mask_loss=0.3, dis_sub_iter=1, gen_sub_iter=1, critic_keep_prob=1.0)
    o32 = OracleAdversarial(P, B, **kw)
    o64 = OracleAdversarial(P, B, dtype=torch.float64, **kw)
    mr, ct = synthetic_images(B, 1234), synthetic_images(B, 4321, 0.3, 0.8)
    rec = {"dis32": [], "dis64": [], "gen32": [], "gen64": [], "scale": [], "dis_err": [], "gen_err": []}
    for s in range(n):
        t0 = time.time()
        d32, d64 = o32.d_step(mr, ct, 1.0), o64.d_step(mr.double(), ct.double(), 1.0)
        g32, g64 = o32.g_step(ct, 1.0), o64.g_step(ct.double(), 1.0)
        sc = 2e-3 * float(d64["mr_cls"].abs().max())
        print("adv step %2d  dis %.6e / %.6e (err/scale %.2e)  gen %.6e / %.6e (err/scale %.2e)  %.1fs" % (
            s, d32["dis_loss"], d64["dis_loss"], abs(d32["dis_loss"] - d64["dis_loss"]) / max(abs(d64["dis_loss"]), sc),
            g32["gen_loss"], g64["gen_loss"], abs(g32["gen_loss"] - g64["gen_loss"]) / max(abs(g64["gen_loss"]), sc),
            time.time() - t0), flush=True)
        for k, v in (("dis32", d32["dis_loss"]), ("dis64", d64["dis_loss"]), ("gen32", g32["gen_loss"]), ("gen64", g64["gen_loss"]),
                     ("scale", sc), ("dis_err", abs(d32["dis_loss"] - d64["dis_loss"]) / max(abs(d64["dis_loss"]), sc)),
                     ("gen_err", abs(g32["gen_loss"] - g64["gen_loss"]) / max(abs(g64["gen_loss"]), sc))):
            rec[k].append(v)
    return rec


if __name__ == "__main__":
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
    which, n = sys.argv[1], int(sys.argv[2])
    rec = (seg if which == "seg" else adv)(n)
    path = os.path.join(ROOT, "tests", "golden", "oracle_trajectory_calibration.json")
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data[which] = rec
    data["how"] = ("scripts/oracle_trajectory_calibration.py: oracle/pnp_graphs.py in fp32 and fp64 from identical initial variables "
                   "(seed 0, stddev 0.05, default BN state) and inputs (B=2, seeds 1234/4321/99), keep_prob 1")
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
