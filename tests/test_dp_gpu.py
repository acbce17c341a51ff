"""Data-parallel parity on 2 GPUs: a 2-rank step == ONE optimizer step on the average of the two ranks'
single-GPU gradients (per-rank batch-norm statistics), i.e. the average of two reference steps at the per-rank batch.
Needs >= 2 CUDA devices (run with `gpurun --gpus 2`); skipped on the 1-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 2


def _build(seed_shift=0):
    import pnp_b200  # noqa: F401
    from pnp_b200 import runtime as rt, adversarial as adv
    from pnp_b200.train_gan import configure
    from oracle.pnp_graphs import OracleAdversarial, init_numpy_params
    rt.set_conv_backend("simt")        # deterministic-order fp32 kernels: the comparison is about the exchange, not rounding
    ws, bns = OracleAdversarial.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    ck, nc, tc = configure("train-gan")
    net = adv.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=ck, network_config=nc, critic_keep_prob=1.0)
    rt.load_state_dict(P)
    tc["dis_sub_iter"] = 1
    tr = adv.Trainer(net, num_cls=5, batch_size=B, opt_kwargs={"learning_rate": 3e-4}, train_config=tc)
    return rt, tr


def _shards():
    from oracle.pnp_graphs import synthetic_images
    return [(synthetic_images(B, 1234 + r), synthetic_images(B, 4321 + r, 0.3, 0.8)) for r in (0, 1)]


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    from pnp_b200 import parallel
    parallel.init_from_env()
    rt, tr = _build()
    assert tr.dp.world == 2
    mr, ct = _shards()[rank]
    tr.d_step(mr.cuda(), ct.cuda(), keep_prob=1.0)
    tr.g_step(ct.cuda(), keep_prob=1.0)
    torch.cuda.synchronize()
    if rank == 0:
        out["d"] = tr.d_arena.theta.cpu().numpy()
        out["g"] = tr.g_arena.theta.cpu().numpy()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_step_equals_average_of_single_gpu_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(29533 + os.getpid() % 1000, out), nprocs=2, join=True)
    # single-process emulation: gradients of each shard from the same initial variables, averaged, one update
    rt, tr = _build()
    grads = []
    theta0 = tr.d_arena.theta.clone()
    for mr, ct in _shards():
        tr.d_arena.theta.copy_(theta0)
        tr.d_step(mr.cuda(), ct.cuda(), keep_prob=1.0, apply=False)
        grads.append(tr.d_arena.grad.clone())
    tr.d_arena.grad.copy_(grads[0] + grads[1])
    tr.dis_optimizer.step(grad_scale=0.5)
    torch.cuda.synchronize()
    ref = tr.d_arena.theta.cpu().numpy()
    err = np.abs(ref - out["d"]).max() / np.abs(ref).max()
    print("  D arena after the 2-rank step vs averaged-gradient emulation: max rel err %.3e" % err)
    assert err <= 1e-5
