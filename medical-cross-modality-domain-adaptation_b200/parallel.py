"""Data parallelism for the hot path: one process per GPU, parameters and optimizer state replicated,
the per-slice minibatch sharded across ranks, and ONE all-reduce (NCCL over NVLink/NVSwitch, or gloo in
the CPU tests) over the flat gradient arena per optimizer step.  The 1/N average is folded into the
optimizer kernel (`grad_scale`), so no extra pass touches the gradients.

The reference is single-GPU (train_segmenter.py:20, train_gan.py:18); parity under DP is defined as:
an N-rank step == the average of N single-GPU reference steps at the per-rank batch (batch-norm
statistics, the CE class weights and the Dice sums stay per rank -- SURVEY 8e).
"""
import os
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    import datetime
    # a rank that stops participating must abort the job quickly instead of hanging the box
    dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=int(os.environ.get("PNP_DIST_TIMEOUT", "180"))))


class DataParallel:
    def __init__(self):
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.on else 1
        self.rank = dist.get_rank() if self.on else 0

    def allreduce(self, grad_arena):
        """sum-all-reduce the flat gradient arena in place; returns the grad_scale (1/world) to hand to
        the optimizer kernel."""
        if self.on:
            dist.all_reduce(grad_arena, op=dist.ReduceOp.SUM)
        return 1.0 / self.world

    def broadcast_params(self, theta_arena):
        """make every replica start from rank 0's parameters"""
        if self.on:
            dist.broadcast(theta_arena, src=0)

    def barrier(self):
        if self.on:
            dist.barrier()

    def broadcast_variables(self, variables):
        """every replica continues from rank 0's values: ALL graph variables (frozen weights and BN moving statistics
        included), after construction and after every restore -- replicas must not depend on identical host seeds"""
        if not self.on:
            return
        seen = set()
        with torch.no_grad():
            for v in variables:
                arena = getattr(v, "_pnp_arena", None)
                t = arena.theta if arena is not None else v
                if id(t) in seen:
                    continue
                seen.add(id(t))
                dist.broadcast(t, src=0)
                if arena is None:
                    v.pnp_version = getattr(v, "pnp_version", 0) + 1
        for v in variables:
            if getattr(v, "_pnp_arena", None) is not None:
                v.pnp_version = getattr(v, "pnp_version", 0) + 1

    def average_moving_stats(self, variables):
        """BN moving statistics are updated from per-rank batches (no SyncBN); a checkpoint stores their mean"""
        if not self.on:
            return
        with torch.no_grad():
            for v in variables:
                if getattr(v, "pnp_kind", "") == "bn_moving":
                    dist.all_reduce(v, op=dist.ReduceOp.SUM)
                    v.mul_(1.0 / self.world)
                    v.pnp_version = getattr(v, "pnp_version", 0) + 1

    def save_checkpoint(self, save_fn):
        """rank 0 alone writes (save_fn must write atomically); every rank waits for the file to be complete"""
        if self.rank == 0:
            save_fn()
        self.barrier()
