#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 256x256x3 slices/sec of the full adversarial step (train_gan.py
--phase train-gan: one discriminator update on B MR + B CT slices incl. the weight clip, then one generator
(DAM) update on B CT slices), B slices per domain per GPU, synthetic data, random-init weights.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU restatement of the reference's TF-1.4 path,
                                                              timed on the host cores -- TF-1.4 itself cannot run here)

Prints ONE JSON line (rank 0).  `value` = whole-job slices/s with inputs resident in HBM; `e2e` = the same metric
through the Trainer API with pinned-host inputs copied every step and the loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# algorithmic conv/FC FLOPs (2*MAC) per unit, SURVEY Appendix A.5 / BASELINE.md section 4
GF_D_STEP_PER_PAIR = 394.90      # D step, lambda_mask > 0, per CT+MR pair
GF_G_STEP_PER_SLICE = 256.16     # G step, per CT slice
METRIC = "slices_per_sec_full_adversarial_step_256x256x3"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "hbm_gbs": d.get("hbm_gbs", 6650.0),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_adversarial(B, backend, seed=0):
    import pnp_b200  # noqa: F401
    from pnp_b200 import runtime as rt, adversarial as adv
    from pnp_b200.train_gan import configure
    rt.set_conv_backend(backend)
    torch.manual_seed(seed)
    rt.manual_seed(1234 + int(os.environ.get("RANK", "0")))
    ck, nc, tc = configure("train-gan")
    net = adv.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=ck, network_config=nc, stddev=0.05, stddev_plain=0.05)
    tc["dis_sub_iter"] = 1           # headline: n_D = 1 discriminator update per generator update
    trainer = adv.Trainer(net, num_cls=5, batch_size=B, opt_kwargs={"learning_rate": 3e-4}, train_config=tc)
    trainer.dp.broadcast_params(trainer.d_arena.theta)
    trainer.dp.broadcast_params(trainer.g_arena.theta)
    return net, trainer


def _host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms use the physical cores instead (logical/2)"""
    n = max(1, (os.cpu_count() or 2) // 2)
    torch.set_num_threads(n)
    return n


def run_ours(a):
    # stdout must carry exactly one JSON line: NCCL prints its version banner to stdout when NCCL_DEBUG >= VERSION, so the variable
    # is cleared (PNP_NCCL_DEBUG re-enables it) and fd 1 points at stderr while the communicator is created (first collective)
    if "PNP_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["PNP_NCCL_DEBUG"]
    else:
        os.environ.pop("NCCL_DEBUG", None)
    from pnp_b200 import parallel, _C
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        parallel.init_from_env()
        if world > 1:
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    B = a.batch
    net, trainer = build_adversarial(B, a.backend)
    from pnp_b200 import functional as F, runtime as rt
    from pnp_b200.data import SyntheticSource

    mr_src = SyntheticSource(B, seed=1234 + rank, pool=3)
    ct_src = SyntheticSource(B, seed=4321 + rank, shift=0.3, scale=0.8, pool=3)
    dev_pool = [(m[0].to(dev), c[0].to(dev)) for m, c in zip(mr_src.pool, ct_src.pool)]

    graphed = False
    if a.graph:
        graphed = trainer.capture_joint_step(dev_pool[0][0], dev_pool[0][1], a.keep_prob)
    launches_per_step = [None]

    def step_resident(i):
        mr, ct = dev_pool[i % len(dev_pool)]
        if graphed and F.PROFILE is None:
            trainer.joint_step(mr, ct, a.keep_prob)
        else:
            trainer.d_step(mr, ct, a.keep_prob)
            trainer.g_step(ct, a.keep_prob)

    def step_e2e(i):
        mr_h, ct_h = mr_src.pool[i % 3][0], ct_src.pool[i % 3][0]
        if graphed:
            d, g = trainer.joint_step(mr_h, ct_h, a.keep_prob)      # pinned host -> static device buffers -> graph replay
        else:
            mr, ct = mr_h.to(dev, non_blocking=True), ct_h.to(dev, non_blocking=True)
            d = trainer.d_step(mr, ct, a.keep_prob)
            g = trainer.g_step(ct, a.keep_prob)
        return trainer.loss_value(d), trainer.loss_value(g)     # .item() reads: device -> host every step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    # kernel launches of one step, counted on an eager step (a graph replay issues the same kernels without host calls)
    l0 = _C.launch_count
    trainer.d_step(dev_pool[0][0], dev_pool[0][1], a.keep_prob)
    trainer.g_step(dev_pool[0][1], a.keep_prob)
    launches_per_step[0] = _C.launch_count - l0
    for i in range(a.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, a.steps)
    launches = launches_per_step[0] * a.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    slices_per_step = 3 * B * world
    value = slices_per_step / (ms_step / 1e3)

    # end-to-end through the public Trainer API with host inputs + loss read-back
    for i in range(min(2, a.warmup)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, a.steps) / a.steps
    e2e = {"value": slices_per_step / (ms_e2e / 1e3), "unit": "slices/s", "h2d_bytes_per_step": 3 * B * 256 * 256 * 3 * 4,
           "d2h_bytes_per_step": 3 * 4, "ms_per_step": ms_e2e}

    # roofline of the dominant kernel (tcgen05 conv): CUDA events around every launch over a few steps
    roof = None
    if rank == 0:
        F.PROFILE = []
    for i in range(min(a.steps, 3)):      # every rank steps (the steps contain the gradient all-reduce); rank 0 records events
        step_resident(i)
    torch.cuda.synchronize()
    if rank == 0:
        recs_all = F.PROFILE
        F.PROFILE = None
        recs = [r_ for r_ in recs_all if not r_[3].startswith("simt:")]      # the roofline is the tcgen05 kernel's
        by_s = {}
        for s_, e_, fl_, tag_, _k in recs_all:
            if tag_.startswith("simt:"):
                c_ = by_s.setdefault((tag_, round(fl_ / 1e9, 3)), [0, 0.0])
                c_[0] += 1
                c_[1] += s_.elapsed_time(e_)
        for (tag_, gf_), (n_, ms_) in sorted(by_s.items(), key=lambda kv: -kv[1][1])[:40]:
            print("[simt] %-24s %9.3f GF x%3d  %8.3f ms  %7.1f TF/s" % (tag_, gf_, n_, ms_, gf_ * n_ / ms_), file=sys.stderr)
        by = {}
        for s_, e_, fl_, tag_, _k in recs:
            k_ = (tag_, round(fl_ / 1e9, 3))
            c_ = by.setdefault(k_, [0, 0.0])
            c_[0] += 1
            c_[1] += s_.elapsed_time(e_)
        top = sorted(by.items(), key=lambda kv: -kv[1][1])[:48]
        for (tag_, gf_), (n_, ms_) in top:
            print("[tc] %-18s %9.3f GF x%3d  %8.3f ms  %7.1f TF/s" % (tag_, gf_, n_, ms_, gf_ * n_ / ms_), file=sys.stderr)
        all_ms = sum(r_[0].elapsed_time(r_[1]) for r_ in recs)
        all_fl = sum(r_[2] for r_ in recs)
        # the dominant kernel = the instantiation with the largest share of the step (agrees with profiles/r1_launches_summary.md)
        per_k = {}
        for s_, e_, fl_, tag_, k_ in recs:
            c_ = per_k.setdefault(k_, [0, 0.0, 0.0])
            c_[0] += 1
            c_[1] += s_.elapsed_time(e_)
            c_[2] += fl_
        for k_, (n_, ms_, fl_) in sorted(per_k.items(), key=lambda kv: -kv[1][1]):
            print("[kern] %-34s x%4d %9.3f ms  %7.1f TF/s" % (k_, n_, ms_, fl_ / ms_ / 1e9), file=sys.stderr)
        dom = max(per_k.items(), key=lambda kv: kv[1][1])[0] if per_k else None
        dom_recs = [r_ for r_ in recs if r_[4] == dom]
        tc_ms = sum(r_[0].elapsed_time(r_[1]) for r_ in dom_recs)
        tc_fl = sum(r_[2] for r_ in dom_recs)
        pk = _peaks()
        nterms = 1 if a.backend == "tc1" else 3
        traffic, traffic_note = None, "no ncu capture committed"
        tj = os.path.join(ROOT, "profiles", "r1_conv_tc_ncu.json")
        if os.path.exists(tj):
            with open(tj) as f:
                nj = json.load(f)
            mine = [l_ for l_ in nj.get("launches", []) if dom and l_["kernel"].endswith(dom)]
            if mine:
                traffic = sum(l_["dram_bytes"] for l_ in mine) / len(mine)
                traffic_note = ("dram__bytes_read+write per launch, mean over the %d %s launches of the committed ncu --set full capture "
                                "(profiles/r1_conv_tc_ncu.csv; different layers than the event-timed mean, same kernel); operands are "
                                "4 B/element (bf16 hi+lo), outputs mostly stay in the 126 MB L2" % (len(mine), dom))
        if recs and tc_ms > 0:
            ach = tc_fl / (tc_ms * 1e-3) / 1e12
            ach_all = all_fl / (all_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "%s (tcgen05.mma kind::f16 + TMA, persistent)" % dom, "achieved": ach, "peak": pk["bf16_tflops"],
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": traffic, "traffic_note": traffic_note, "peak_source": pk["source"],
                    "launches_per_step": len(dom_recs) / min(a.steps, 3), "kernel_ms_per_step": tc_ms / min(a.steps, 3),
                    "share_of_step": (tc_ms / min(a.steps, 3)) / ms_step, "mma_terms": nterms,
                    "issued_frac": nterms * ach / pk["bf16_tflops"],
                    "all_tcgen05_convs": {"achieved": ach_all, "frac": ach_all / pk["bf16_tflops"], "issued_frac": nterms * ach_all / pk["bf16_tflops"],
                                          "launches_per_step": len(recs) / min(a.steps, 3), "kernel_ms_per_step": all_ms / min(a.steps, 3),
                                          "share_of_step": (all_ms / min(a.steps, 3)) / ms_step},
                    "note": "achieved = algorithmic 2*M*N*K per launch / event time; the fp32-grade path issues mma_terms bf16 MMAs per "
                            "algorithmic MAC, so tensor-pipe occupancy ~ issued_frac"}
        else:
            roof = {"bound": "tensor", "achieved": 0.0, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": 0.0, "traffic": None,
                    "note": "no tcgen05 launches recorded (backend=%s)" % a.backend}

    if a.profile and rank == 0 and world == 1:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                step_resident(i)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:45]
        tot = sum(e.device_time_total for e in prof.key_averages())
        for e in rows:
            print("[prof] %-110s x%4d %9.3f ms %5.1f%%" % (e.key.replace("(anonymous namespace)::", "")[:110], e.count, e.device_time_total / 1e3 / 2, 100.0 * e.device_time_total / tot),
                  file=sys.stderr)
        print("[prof] total device time per step %.3f ms" % (tot / 1e3 / 2), file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline_sample(2, 1)

    if rank == 0:
        gf_step = B * (GF_D_STEP_PER_PAIR + GF_G_STEP_PER_SLICE) * world
        out = {
            "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tcgen05 bf16 hi/lo split x%d, fp32 accumulate)" % (1 if a.backend == "tc1" else 3), "data": "synthetic",
            "config": {"workload": "train_gan.py --phase train-gan joint step: 1 D update (B MR + B CT, +clip) + 1 G update (B CT); "
                                   "BASELINE configs[3] at N GPUs", "batch_per_gpu_per_domain": B, "slices_per_step": slices_per_step,
                       "keep_prob": a.keep_prob, "conv_backend": a.backend, "parallelism": "dp%d" % world, "cuda_graph": bool(graphed),
                       "l2": "per-step working set (activations of %d slices, GBs) exceeds the 126 MB L2; no explicit flush" % (3 * B)},
            "conv_tflops_algorithmic": gf_step / ms_step / 1e3,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        # teardown must never hang the launcher: drop the captured graph (it pins NCCL work) before the last barrier, and
        # leave through a watchdog-protected hard exit
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        trainer._graph = None
        trainer._graph_out = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        try:
            dist.barrier()
        except Exception:
            pass
        os._exit(0)


def cpu_baseline_sample(B, reps):
    """the oracle's joint adversarial step (same math, torch-CPU/oneDNN) on the host cores -- bounded sample"""
    from oracle.pnp_graphs import OracleAdversarial, init_numpy_params, synthetic_images
    # physical cores: forcing os.cpu_count() logical threads oversubscribes oneDNN on the GPU host and is >10x slower
    _host_threads()
    ws, bns = OracleAdversarial.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    o = OracleAdversarial(P, B, lambda_mask_loss=0.3, dis_sub_iter=1, gen_sub_iter=1, critic_keep_prob=0.75)
    mr, ct = synthetic_images(B, 1234), synthetic_images(B, 4321, 0.3, 0.8)
    times = []
    for _ in range(reps):
        t0 = time.time()
        o.d_step(mr, ct, 0.75)
        o.g_step(ct, 0.75)
        times.append(time.time() - t0)
    t = sorted(times)[len(times) // 2]
    return {"value": 3 * B / t, "unit": "slices/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d joint D+G step(s) at B=%d per domain (%.1f s each); TF-1.4 semantics restated on torch-CPU, TF itself "
                      "cannot run in this image" % (reps, B, t)}


def run_reference(a):
    """--impl reference: the reference's own CPU path = the oracle port (oracle/pnp_graphs.py), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.pnp_graphs import OracleAdversarial, init_numpy_params, synthetic_images
    _host_threads()
    B = 1
    ws, bns = OracleAdversarial.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    o = OracleAdversarial(P, B, lambda_mask_loss=0.3, dis_sub_iter=1, gen_sub_iter=1, critic_keep_prob=0.75)
    mr, ct = synthetic_images(B, 1234), synthetic_images(B, 4321, 0.3, 0.8)

    def step():
        o.d_step(mr, ct, 0.75)
        o.g_step(ct, 0.75)
    for _ in range(a.warmup):
        step()
    t0 = time.time()
    for _ in range(a.steps):
        step()
    dt = (time.time() - t0) / a.steps
    v = 3 * B / dt
    sample = "each step = one joint D+G step at B=1 per domain on the host cores (bounded sample of the B=%d/GPU workload)" % a.batch
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "slices/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "train_gan.py --phase train-gan joint step (CPU restatement of the TF-1.4 reference path)",
                      "batch_per_gpu_per_domain": a.batch, "sample": sample},
           "cpu_baseline": {"value": v, "unit": "slices/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="slices per domain per GPU (BASELINE configs[3]: 8/GPU)")
    ap.add_argument("--keep-prob", type=float, default=0.75)
    ap.add_argument("--backend", default="auto", choices=["auto", "simt", "tc3", "tc1"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True, help="replay the step as one CUDA graph (default)")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--profile", action="store_true", help="print a per-kernel device-time table (torch profiler) to stderr")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(a)


if __name__ == "__main__":
    main()
