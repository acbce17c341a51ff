"""Flat parameter / gradient arenas and the fused optimizers of the hot path.

All variables of one optimizer live in ONE contiguous fp32 arena (each padded to 1024 floats), their
gradients in a second arena of the same layout; `var.data` / `var.grad` are views.  One kernel launch
updates everything (TF Adam, source_segmenter.py:378; TF RMSProp + WGAN weight clip,
adversarial.py:643-654), folding in the L2-regulariser gradient (wd * theta) and the 1/N data-parallel
average; one NCCL all-reduce over the gradient arena is all the communication DP needs (parallel.py).
"""
import torch

from . import runtime as rt
from ._C import call, ptr

CHUNK = 1024


class Arena:
    def __init__(self, variables):
        variables = list(variables)
        if not variables:
            raise ValueError("empty variable list")
        dev = variables[0].device
        self.vars = variables
        self.offsets = []
        off = 0
        seg_of_chunk = []
        for i, v in enumerate(variables):
            if getattr(v, "_pnp_arena", None) is not None:
                raise ValueError("variable %s already belongs to an arena" % getattr(v, "pnp_name", "?"))
            n = v.numel()
            padded = -(-n // CHUNK) * CHUNK
            self.offsets.append((off, n))
            seg_of_chunk += [i] * (padded // CHUNK)
            off += padded
        self.total = off
        self.theta = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.chunk_seg = torch.tensor(seg_of_chunk, dtype=torch.int32, device=dev)
        with torch.no_grad():
            for v, (o, n) in zip(variables, self.offsets):
                view = self.theta[o:o + n].view(v.shape)
                view.copy_(v)
                v.data = view
                v.grad = self.grad[o:o + n].view(v.shape)
                v._pnp_arena = self

    def zero_grad(self):
        call("pnp_fill", ptr(self.grad), 0.0, self.total, rt.stream())

    def bump_versions(self):
        for v in self.vars:
            v.pnp_version = getattr(v, "pnp_version", 0) + 1

    def seg_table(self, values):
        return torch.tensor([float(x) for x in values], dtype=torch.float32, device=self.theta.device)


def _slot_views(arena, flat):
    return [flat[o:o + n].view(v.shape) for v, (o, n) in zip(arena.vars, arena.offsets)]


def _export_slots(arena, slots):
    """{'<variable name>/<slot>': numpy} -- tf.train.Saver stores optimizer slots under these names"""
    out = {}
    for slot_name, flat in slots.items():
        for v, view in zip(arena.vars, _slot_views(arena, flat)):
            out["%s/%s" % (v.pnp_name, slot_name)] = view.detach().cpu().numpy().copy()
    return out


def _import_slots(arena, slots, d):
    """load what is stored; returns the number of slot tensors found"""
    found = 0
    with torch.no_grad():
        for slot_name, flat in slots.items():
            for v, view in zip(arena.vars, _slot_views(arena, flat)):
                k = "%s/%s" % (v.pnp_name, slot_name)
                if k in d:
                    view.copy_(torch.as_tensor(d[k], dtype=view.dtype).reshape(view.shape))
                    found += 1
    return found


class Adam:
    """tf.train.AdamOptimizer(learning_rate, beta1=.9, beta2=.999, epsilon=1e-8) over an Arena."""

    def __init__(self, arena, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=None):
        self.arena = arena
        self.b1, self.b2, self.eps = beta1, beta2, eps
        dev = arena.theta.device
        self.m = torch.zeros_like(arena.theta)
        self.v = torch.zeros_like(arena.theta)
        self.state = torch.tensor([1.0, 1.0, float(lr), 0.0], dtype=torch.float64, device=dev)
        self.seg_wd = arena.seg_table(weight_decay if weight_decay is not None else [0.0] * len(arena.vars))
        self.t = 0

    def set_lr(self, lr):
        self.state[2] = float(lr)

    def get_lr(self):
        return float(self.state[2].item())

    def slot_state(self):
        """tf.train.AdamOptimizer slots '<var>/Adam' (m), '<var>/Adam_1' (v) + the beta-power accumulators"""
        d = _export_slots(self.arena, {"Adam": self.m, "Adam_1": self.v})
        st = self.state.detach().cpu().numpy()
        d["beta1_power"], d["beta2_power"] = st[0].astype("float32"), st[1].astype("float32")
        return d

    def load_slot_state(self, d):
        n = _import_slots(self.arena, {"Adam": self.m, "Adam_1": self.v}, d)
        if "beta1_power" in d and "beta2_power" in d:
            self.state[0] = float(d["beta1_power"])
            self.state[1] = float(d["beta2_power"])
        return n

    def step(self, grad_scale=1.0):
        a = self.arena
        self.t += 1
        call("pnp_adam_advance", ptr(self.state), self.b1, self.b2, rt.stream())
        call("pnp_adam_step", ptr(a.theta), ptr(a.grad), ptr(self.m), ptr(self.v), a.total, ptr(a.chunk_seg), ptr(self.seg_wd),
             ptr(self.state), self.b1, self.b2, self.eps, float(grad_scale), rt.stream())
        a.bump_versions()


class RMSProp:
    """tf.train.RMSPropOptimizer(learning_rate, decay=.9, momentum=0, epsilon=1e-10) over an Arena,
    `ms` initialised to ONE, epsilon inside the sqrt; optional per-variable clip (WGAN clip_op)."""

    def __init__(self, arena, lr=3e-4, decay=0.9, momentum=0.0, eps=1e-10, weight_decay=None, clip=None):
        self.arena = arena
        self.decay, self.momentum, self.eps = decay, momentum, eps
        dev = arena.theta.device
        self.ms = torch.ones_like(arena.theta)
        self.mom = torch.zeros_like(arena.theta)
        self.lr_t = torch.tensor([float(lr)], dtype=torch.float32, device=dev)
        self.seg_wd = arena.seg_table(weight_decay if weight_decay is not None else [0.0] * len(arena.vars))
        self.seg_clip = arena.seg_table(clip if clip is not None else [0.0] * len(arena.vars))

    def set_lr(self, lr):
        self.lr_t[0] = float(lr)

    def get_lr(self):
        return float(self.lr_t.item())

    def slot_state(self):
        """tf.train.RMSPropOptimizer slots '<var>/RMSProp' (ms), '<var>/RMSProp_1' (momentum)"""
        return _export_slots(self.arena, {"RMSProp": self.ms, "RMSProp_1": self.mom})

    def load_slot_state(self, d):
        return _import_slots(self.arena, {"RMSProp": self.ms, "RMSProp_1": self.mom}, d)

    def set_weight_decay(self, values):
        self.seg_wd.copy_(self.arena.seg_table(values))

    def step(self, grad_scale=1.0):
        a = self.arena
        call("pnp_rmsprop_step", ptr(a.theta), ptr(a.grad), ptr(self.ms), ptr(self.mom), a.total, ptr(a.chunk_seg), ptr(self.seg_wd),
             ptr(self.seg_clip), ptr(self.lr_t), self.decay, self.momentum, self.eps, float(grad_scale), rt.stream())
        a.bump_versions()


class Momentum:
    """tf.train.MomentumOptimizer(learning_rate=exponential_decay(lr, global_step, decay_steps, decay_rate, staircase=True),
    momentum) over an Arena -- the source segmenter's `optimizer="momentum"` branch (source_segmenter.py:360-372; defaults
    learning_rate 0.2, decay_rate 0.95, momentum 0.2)."""

    def __init__(self, arena, lr=0.2, decay_rate=0.95, momentum=0.2, decay_steps=100, weight_decay=None):
        self.arena = arena
        self.lr0, self.decay_rate, self.momentum, self.decay_steps = float(lr), float(decay_rate), float(momentum), int(decay_steps)
        dev = arena.theta.device
        self.accum = torch.zeros_like(arena.theta)
        self.lr_t = torch.tensor([float(lr)], dtype=torch.float32, device=dev)
        self.seg_wd = arena.seg_table(weight_decay if weight_decay is not None else [0.0] * len(arena.vars))
        self.global_step = 0

    def current_lr(self):
        return self.lr0 * self.decay_rate ** (self.global_step // max(1, self.decay_steps))

    def set_lr(self, lr):
        self.lr0 = float(lr)

    def get_lr(self):
        return self.current_lr()

    def slot_state(self):
        return _export_slots(self.arena, {"Momentum": self.accum})

    def load_slot_state(self, d):
        return _import_slots(self.arena, {"Momentum": self.accum}, d)

    def step(self, grad_scale=1.0):
        a = self.arena
        self.lr_t[0] = self.current_lr()                 # staircase decay is evaluated at the step's global_step, as TF does
        call("pnp_momentum_step", ptr(a.theta), ptr(a.grad), ptr(self.accum), a.total, ptr(a.chunk_seg), ptr(self.seg_wd),
             ptr(self.lr_t), self.momentum, float(grad_scale), rt.stream())
        self.global_step += 1
        a.bump_versions()
