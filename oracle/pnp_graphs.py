"""
ORACLE (test infrastructure, NOT product code) -- CPU restatement of the two reference graphs and
their training steps, built on oracle/tf14_torch.py:

  * OracleSegmenter   <- source_segmenter.py:48-273 (Full_DRN graph + losses), :357-381 (Adam)
  * OracleAdversarial <- adversarial.py:44-501 (two-stream fronts, shared back half, feature
                         discriminator, mask critic, WGAN losses), :633-656 (RMSProp x2 + clip),
                         :840-882 (feed conventions of the D step and the G step)

PARITY STATUS (see oracle/tf14_numpy.py and DESIGN.md section 2): TF-1.4 cannot run in this image and the reference's own
modules do not run as they are (source_segmenter.py:611 syntax error, adversarial.py:102 attribute typo), so the op
NUMERICS here are an unpinned restatement.  The loss arithmetic (dis_losses / gen_losses) is pinned to the reference's
adversarial.py:445-476 executed numerically (tests/test_reference_graph_trace.py); the architecture, variable tables,
optimizer wiring and step feeds of the reference -- obtained by executing its graph code under a recording tensorflow
shim -- are pinned both on the PRODUCT side and for the graphs of this file (layer lists, variable layouts, L2 lists:
test_oracle_graphs_match_the_reference_trace / test_oracle_segmenter_matches_the_reference_trace).

Variables are keyed by the TF variable names the reference would create (the checkpoint naming
contract of lists/half_zip_*_vars, lists/*_bn_list), so the same numpy dict initialises both this
oracle and the CUDA product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import numpy as np
import torch

from . import tf14_torch as T
from .tf14_numpy import truncated_normal

FB = 16  # feature_base (source_segmenter.py:76, adversarial.py:84)

# (group index, [(kind, cin, cout)]) -- source_segmenter.py:91-161 == adversarial.py:130-269
FRONT_GROUPS = [
    (1, [("conv", 3, FB), ("res", FB, FB), ("pool",)]),
    (2, [("res", FB, 2 * FB), ("pool",)]),
    (3, [("res", 2 * FB, 4 * FB), ("res", 4 * FB, 4 * FB), ("pool",)]),
    (4, [("res", 4 * FB, 8 * FB), ("res", 8 * FB, 8 * FB)]),
    (5, [("res", 8 * FB, 16 * FB), ("res", 16 * FB, 16 * FB)]),
    (6, [("res", 16 * FB, 16 * FB), ("res", 16 * FB, 16 * FB)]),
]
# source_segmenter.py:163-209 == adversarial.py:273-318
BACK_GROUPS = [
    (7, [("res", 16 * FB, 32 * FB), ("res", 32 * FB, 32 * FB)]),
    (8, [("dr", 32 * FB, 32 * FB), ("dr", 32 * FB, 32 * FB)]),
    (9, [("cbr", 32 * FB, 32 * FB), ("cbr", 32 * FB, 32 * FB)]),
]


def _vname(scope, j):
    return "%s/Variable" % scope if j == 0 else "%s/Variable_%d" % (scope, j)


def _weights_of_groups(groups, scope_fmt):
    """yield (name, shape) for every conv weight in creation order."""
    for gi, ops in groups:
        scope = scope_fmt % gi
        j = 0
        for op in ops:
            if op[0] == "pool":
                continue
            if op[0] in ("conv", "cbr"):
                yield _vname(scope, j), (3, 3, op[1], op[2])
                j += 1
            else:  # res / dr: two 3x3 convs
                yield _vname(scope, j), (3, 3, op[1], op[2])
                yield _vname(scope, j + 1), (3, 3, op[2], op[2])
                j += 2


def seg_tail_weights(num_cls=5):
    return [("group_10/Variable", (3, 3, 32 * FB, 64 * num_cls * 8)),
            ("output/Variable", (5, 5, num_cls * 8, num_cls))]


# ---- BN scope naming -------------------------------------------------------------------------
def _bn_scopes_front(prefix_kind):
    """BN scope names (with channel counts) in creation order for groups 1-6.
    prefix_kind: 'anon' (source_segmenter: BatchNorm, BatchNorm_1, ...), 'pred' (adversarial MR
    path: group_k/pred_k_b_{1,2}), 'adapt' (adversarial CT path, adversarial.py:206-267: note the
    irregular scope strings 'adapt_1', 'adapt_2' for groups 1-2 and 'adapt_k_b' afterwards)."""
    out = []
    for gi, ops in FRONT_GROUPS:
        b = 0
        for op in ops:
            if op[0] != "res":
                continue
            b += 1
            if prefix_kind == "pred":
                base = "group_%d/pred_%d_%d" % (gi, gi, b)
            elif prefix_kind == "adapt":
                base = "adapt_%d/adapt_%d" % (gi, gi) if gi <= 2 else "adapt_%d/adapt_%d_%d" % (gi, gi, b)
            else:
                base = None
            out.append((base, op[2]))
    return out


def _bn_scopes_back_pred():
    out = []
    for gi, ops in BACK_GROUPS:
        b = 0
        for op in ops:
            b += 1
            base = "group_%d/pred_%d_%d" % (gi, gi, b)
            out.append((base, op[0], op[2]))
    return out


class ParamStore:
    """name -> tensor for conv/FC weights, name -> BNState for batch-norm scopes."""

    def __init__(self, dtype=torch.float32):
        self.dtype = dtype
        self.w = {}
        self.bn = {}

    def add_w(self, name, arr):
        self.w[name] = torch.tensor(np.asarray(arr), dtype=self.dtype).requires_grad_(True)

    def add_bn(self, scope, C):
        self.bn[scope] = T.BNState(C, self.dtype)

    def load_numpy(self, d):
        """d: flat dict of numpy arrays keyed by TF names ('.../Variable', '.../beta', ...)."""
        with torch.no_grad():
            for k, v in d.items():
                t = torch.tensor(np.asarray(v), dtype=self.dtype)
                if k in self.w:
                    self.w[k].copy_(t)
                else:
                    scope, leaf = k.rsplit("/", 1)
                    bn = self.bn[scope]
                    if leaf == "gamma":
                        bn.gamma.copy_(t)
                    elif leaf == "beta":
                        bn.beta.copy_(t)
                    elif leaf == "moving_mean":
                        bn.moving_mean = t.clone()
                    elif leaf == "moving_variance":
                        bn.moving_var = t.clone()
                    else:
                        raise KeyError(k)

    def to_numpy(self):
        d = {k: v.detach().numpy().copy() for k, v in self.w.items()}
        for s, bn in self.bn.items():
            d[s + "/gamma"] = bn.gamma.detach().numpy().copy()
            d[s + "/beta"] = bn.beta.detach().numpy().copy()
            d[s + "/moving_mean"] = bn.moving_mean.numpy().copy()
            d[s + "/moving_variance"] = bn.moving_var.numpy().copy()
        return d


def init_numpy_params(names_shapes, bn_scopes, seed, stddev):
    """Seeded truncated-normal weights (layers.py:47-55) + default BN init (beta 0, gamma 1, moving
    mean 0, moving variance 1)."""
    rng = np.random.RandomState(seed)
    d = {}
    for name, shape in names_shapes:
        d[name] = truncated_normal(rng, shape, stddev)
    for scope, C in bn_scopes:
        d[scope + "/gamma"] = np.ones(C, np.float32)
        d[scope + "/beta"] = np.zeros(C, np.float32)
        d[scope + "/moving_mean"] = np.zeros(C, np.float32)
        d[scope + "/moving_variance"] = np.ones(C, np.float32)
    return d


# ---- shared sub-graphs ------------------------------------------------------------------------
def run_front(ps, x, scope_fmt, bn_names, keep_prob, is_train):
    """groups 1-6 (source_segmenter.py:91-161 / adversarial.py:130-269).  bn_names: list of BN
    scope base names per residual block in order (each block uses base+'_1', base+'_2').
    Returns (c4_2, c6_2)."""
    h = x
    bi = 0
    taps = {}
    for gi, ops in FRONT_GROUPS:
        scope = scope_fmt % gi
        j = 0
        for op in ops:
            if op[0] == "conv":
                h = T.conv2d(h, ps.w[_vname(scope, j)], keep_prob)
                j += 1
            elif op[0] == "res":
                base = bn_names[bi]
                bi += 1
                h = T.residual_block(h, ps.w[_vname(scope, j)], ps.w[_vname(scope, j + 1)], keep_prob,
                                     ps.bn[base + "_1"], ps.bn[base + "_2"], inc_dim=(op[1] != op[2]),
                                     is_train=is_train, leak=True)
                j += 2
            elif op[0] == "pool":
                h = T.max_pool2d(h, 2)
        taps[gi] = h
    return taps[4], taps[6]


def run_back(ps, c6, bn_names, keep_prob, is_train, batch_size, num_cls=5):
    """groups 7-10 + output (source_segmenter.py:163-209 / adversarial.py:273-318).
    bn_names: 4 block bases (g7 x2, g8 x2) then 2 plain scopes (g9).  Returns (c9_2, b8, b7, logits)."""
    w = ps.w
    h = T.residual_block(c6, w["group_7/Variable"], w["group_7/Variable_1"], keep_prob,
                         ps.bn[bn_names[0] + "_1"], ps.bn[bn_names[0] + "_2"], inc_dim=True, is_train=is_train, leak=True)
    b7 = T.residual_block(h, w["group_7/Variable_2"], w["group_7/Variable_3"], keep_prob,
                          ps.bn[bn_names[1] + "_1"], ps.bn[bn_names[1] + "_2"], is_train=is_train, leak=True)
    h = T.DR_block(b7, w["group_8/Variable"], w["group_8/Variable_1"], 2, keep_prob,
                   ps.bn[bn_names[2] + "_1"], ps.bn[bn_names[2] + "_2"], is_train=is_train, leak=True)
    b8 = T.DR_block(h, w["group_8/Variable_2"], w["group_8/Variable_3"], 2, keep_prob,
                    ps.bn[bn_names[3] + "_1"], ps.bn[bn_names[3] + "_2"], is_train=is_train, leak=True)
    h = T.conv_bn_relu2d(b8, w["group_9/Variable"], keep_prob, ps.bn[bn_names[4]], is_train=is_train, leak=True)
    c9 = T.conv_bn_relu2d(h, w["group_9/Variable_1"], keep_prob, ps.bn[bn_names[5]], is_train=is_train, leak=True)
    c10 = T.conv2d(c9, w["group_10/Variable"], keep_prob, padding="SYMMETRIC")
    flat = T.PS(c10, 8, num_cls * 8, batch_size)
    logits = T.conv2d(flat, w["output/Variable"], 1.0, padding="SYMMETRIC")
    return c9, b8, b7, logits


# ==============================================================================================
# source-only segmenter
# ==============================================================================================
class OracleSegmenter:
    """source_segmenter.Full_DRN + the Adam train step of source_segmenter.Trainer."""

    @staticmethod
    def layout(num_cls=5):
        ws = list(_weights_of_groups(FRONT_GROUPS + BACK_GROUPS, "group_%d")) + seg_tail_weights(num_cls)
        bns = []
        k = 0
        for gi, ops in FRONT_GROUPS + BACK_GROUPS:
            for op in ops:
                if op[0] in ("res", "dr"):
                    n = 2
                elif op[0] == "cbr":
                    n = 1
                else:
                    n = 0
                for _ in range(n):
                    bns.append(("BatchNorm" if k == 0 else "BatchNorm_%d" % k, op[2]))
                    k += 1
        return ws, bns

    def __init__(self, params, batch_size, num_cls=5, dtype=torch.float32,
                 miu_dice=1.0, miu_cross=1.0, regularizer=1e-4, lr=1e-3):
        self.batch_size, self.num_cls = batch_size, num_cls
        self.ps = ParamStore(dtype)
        ws, bns = self.layout(num_cls)
        for n, s in ws:
            self.ps.add_w(n, np.zeros(s, np.float32))
        for n, c in bns:
            self.ps.add_bn(n, c)
        self.ps.load_numpy(params)
        self.miu_dice, self.miu_cross, self.reg = miu_dice, miu_cross, regularizer
        # anonymous BN scopes: residual block k uses BatchNorm_{2k}, BatchNorm_{2k+1}
        self._bn_order = [n for n, _ in bns]
        # conv_weights list with the reference's quirk: wr4_4 twice, wr4_3 never (source_segmenter.py:132-135)
        names = [n for n, _ in ws]
        self.l2_names = [n for n in names if n != "group_4/Variable_2"] + ["group_4/Variable_3"]
        trainables = [self.ps.w[n] for n in names]
        for n in self._bn_order:
            trainables += [self.ps.bn[n].gamma, self.ps.bn[n].beta]
        self.trainables = trainables
        self.opt = T.TFAdam(trainables, lr=lr)

    class _AnonBN:
        """maps block base names 'k' -> BatchNorm_{..} in creation order."""

    def forward(self, x, keep_prob=1.0, bn_train=True):
        ps = self.ps
        # build a view of ps.bn keyed by synthetic block names so run_front/run_back can be shared
        order = self._bn_order
        alias = {}
        names_front, k = [], 0
        for _ in range(10):  # 10 residual blocks in groups 1-6
            base = "blk%d" % len(names_front)
            alias[base + "_1"], alias[base + "_2"] = ps.bn[order[k]], ps.bn[order[k + 1]]
            names_front.append(base)
            k += 2
        names_back = []
        for _ in range(4):
            base = "bblk%d" % len(names_back)
            alias[base + "_1"], alias[base + "_2"] = ps.bn[order[k]], ps.bn[order[k + 1]]
            names_back.append(base)
            k += 2
        for i in range(2):
            alias["g9_%d" % i] = ps.bn[order[k]]
            names_back.append("g9_%d" % i)
            k += 1
        saved = ps.bn
        ps.bn = alias
        try:
            c4, c6 = run_front(ps, x, "group_%d", names_front, keep_prob, bn_train)
            c9, b8, b7, logits = run_back(ps, c6, names_back, keep_prob, bn_train, self.batch_size, self.num_cls)
        finally:
            ps.bn = saved
        return {"logits": logits, "c4_2": c4, "c6_2": c6, "b7": b7, "b8": b8, "c9_2": c9}

    def losses(self, logits, y):
        """source_segmenter.py:211-239."""
        wce = T.softmax_weighted_loss(logits, y)
        dice = T.dice_loss(logits, y)
        cost = self.miu_cross * wce + self.miu_dice * dice
        reg = self.reg * sum(T.l2_loss(self.ps.w[n]) for n in self.l2_names)
        return cost, reg, wce, dice

    def evaluate(self, x, y):
        with torch.no_grad():
            out = self.forward(x, 1.0, False)
            pred = T.pixel_wise_softmax_2(out["logits"])
            compact = pred.argmax(3)
            d, arr = T.dice_eval(compact, y, self.num_cls)
        return float(d), [float(a) for a in arr], compact

    def train_step(self, x, y, keep_prob=1.0):
        """One sess.run(optimizer) of source_segmenter.py:484-489 (BN switches True)."""
        out = self.forward(x, keep_prob, True)
        cost, reg, wce, dice = self.losses(out["logits"], y)
        grads = torch.autograd.grad(cost + reg, self.trainables, allow_unused=True)
        self.opt.step(grads)
        return {"cost": float(cost), "reg": float(reg), "wce": float(wce), "dice": float(dice), "grads": grads}


# ==============================================================================================
# adversarial graph
# ==============================================================================================
def disc_input(c4, c6, b7, c9, logits, B):
    """adversarial.py:324-335: the 32-channel discriminator input [PS(c4, 2 groups) tiled x3 | PS(c6, 4) | PS(b7, 8) | PS(c9, 8) | logits |
    float(argmax logits)]; pinned numerically to the executed reference lines (tests/golden/make_reference_disc_input_vectors.py)"""
    f4 = T.PS(c4, 8, 2, B).repeat(1, 1, 1, 3)
    f6 = T.PS(c6, 8, 4, B)
    f7 = T.PS(b7, 8, 8, B)
    f9 = T.PS(c9, 8, 8, B)
    am = logits.argmax(3).to(logits.dtype).unsqueeze(3)
    return torch.cat([f4, f6, f7, f9, logits, am], dim=3)


# (scope, weight shapes, [bn scopes]) for the feature discriminator, adversarial.py:337-398
CLS_BLOCKS = [
    # name, cin, cout, inc, down_k, down_stride
    ("cls_1", 2 * FB, 4 * FB, True, 3, 2),
    ("cls_2", 4 * FB, 8 * FB, True, 5, 2),
    ("cls_3", 8 * FB, 16 * FB, True, 3, 2),
    ("cls_4", 16 * FB, 32 * FB, True, 3, 2),
    ("cls_5", 32 * FB, 32 * FB, False, 5, 4),
]


class OracleAdversarial:
    """adversarial.Full_DRN + the D step / G step of adversarial.Trainer."""

    @staticmethod
    def layout(num_cls=5):
        ws = list(_weights_of_groups(FRONT_GROUPS, "group_%d"))
        ws += list(_weights_of_groups(BACK_GROUPS, "group_%d")) + seg_tail_weights(num_cls)
        ws += list(_weights_of_groups(FRONT_GROUPS, "adapt_%d"))
        bns = []
        for base, C in _bn_scopes_front("pred"):
            bns += [(base + "_1", C), (base + "_2", C)]
        for base, kind, C in _bn_scopes_back_pred():
            bns += [(base, C)] if kind == "cbr" else [(base + "_1", C), (base + "_2", C)]
        for base, C in _bn_scopes_front("adapt"):
            bns += [(base + "_1", C), (base + "_2", C)]
        # feature discriminator
        for name, cin, cout, inc, dk, ds in CLS_BLOCKS:
            s = "cls_scope/" + name
            ws += [(s + "/Variable", (3, 3, cin, cout)), (s + "/Variable_1", (3, 3, cout, cout)),
                   (s + "/Variable_2", (dk, dk, cout, cout))]
            bns += [(s + "/" + name + "_1", cout), (s + "/" + name + "_2", cout), (s + "/" + name + "_3", cout)]
        ws += [("cls_scope/cls_6/Variable", (3, 3, 32 * FB, 32 * FB)), ("cls_scope/cls_out/Variable", (32 * FB * 4, 1))]
        bns += [("cls_scope/cls_6/cls_6", 32 * FB)]
        # mask critic, adversarial.py:402-443
        m = "mask_cls_scope/"
        ws += [(m + "mask_cls_1/Variable", (3, 3, num_cls, FB)),
               (m + "mask_cls_2/Variable", (3, 3, FB, FB)), (m + "mask_cls_2/Variable_1", (3, 3, FB, FB)),
               (m + "mask_cls_2/Variable_2", (5, 5, FB, 2 * FB)),
               (m + "mask_cls_3/Variable", (3, 3, 2 * FB, 4 * FB)), (m + "mask_cls_3/Variable_1", (3, 3, 4 * FB, 4 * FB)),
               (m + "mask_cls_3/Variable_2", (5, 5, 4 * FB, 8 * FB)),
               (m + "mask_cls_4/Variable", (5, 5, 8 * FB, 16 * FB)),
               (m + "m_cls_out/Variable", (16 * FB * 4, 1))]
        bns += [(m + "mask_cls_1/mask_cls_1", FB),
                (m + "mask_cls_2/m_cls_2_1", FB), (m + "mask_cls_2/m_cls_2_2", FB), (m + "mask_cls_2/m_cls_2_3", 2 * FB),
                (m + "mask_cls_3/m_cls_3_1", 4 * FB), (m + "mask_cls_3/m_cls_3_2", 4 * FB), (m + "mask_cls_3/m_cls_3_3", 8 * FB),
                (m + "mask_cls_4/m_cls_4", 16 * FB)]
        return ws, bns

    def __init__(self, params, batch_size, num_cls=5, dtype=torch.float32, miu_dis=0.002, miu_gen=0.002,
                 lambda_mask_loss=0.3, gan_regularizer=1e-4, regularizer=1e-4, lr=3e-4,
                 dis_sub_iter=20, gen_sub_iter=1, critic_keep_prob=0.75):
        self.batch_size, self.num_cls = batch_size, num_cls
        self.ps = ParamStore(dtype)
        ws, bns = self.layout(num_cls)
        for n, s in ws:
            self.ps.add_w(n, np.zeros(s, np.float32))
        for n, c in bns:
            self.ps.add_bn(n, c)
        self.ps.load_numpy(params)
        self.miu_dis, self.miu_gen, self.lam = miu_dis, miu_gen, lambda_mask_loss
        self.gan_reg, self.reg = gan_regularizer, regularizer
        self.dis_sub_iter, self.gen_sub_iter = dis_sub_iter, gen_sub_iter
        # reference hard-wires 0.75 (adversarial.py:320,402); parity runs override to 1.0
        self.critic_keep_prob = critic_keep_prob
        self.pred_front = [b for b, _ in _bn_scopes_front("pred")]
        self.adapt_front = [b for b, _ in _bn_scopes_front("adapt")]
        self.pred_back = [b for b, _, _ in _bn_scopes_back_pred()]
        wnames = [n for n, _ in ws]
        bnames = [n for n, _ in bns]
        # adversarial.py:478-501 -- membership by substring of the variable name
        self.cls_w = [n for n in wnames if "cls" in n]
        self.cls_bn = [n for n in bnames if "cls" in n]
        self.adapt_w = [n for n in wnames if "cls" not in n and "adapt" in n]
        self.adapt_bn = [n for n in bnames if "cls" not in n and "adapt" in n]
        self.d_weights = [n for n in self.cls_w if n.startswith("cls_scope")]
        self.m_weights = [n for n in self.cls_w if n.startswith("mask_cls_scope")]
        self.d_trainables = [self.ps.w[n] for n in self.cls_w]
        for n in self.cls_bn:
            self.d_trainables += [self.ps.bn[n].gamma, self.ps.bn[n].beta]
        self.g_trainables = [self.ps.w[n] for n in self.adapt_w]
        for n in self.adapt_bn:
            self.g_trainables += [self.ps.bn[n].gamma, self.ps.bn[n].beta]
        self.d_opt = T.TFRMSProp(self.d_trainables, lr=lr)
        self.g_opt = T.TFRMSProp(self.g_trainables, lr=lr)

    # ---- sub-graphs ---------------------------------------------------------------------------
    def classifier(self, c4, c6, b7, c9, logits):
        """adversarial.py:320-400.  BN always batch statistics; dropout always critic_keep_prob."""
        ps, B, kp = self.ps, self.batch_size, self.critic_keep_prob
        h = disc_input(c4, c6, b7, c9, logits, B)
        d_input = h
        for name, cin, cout, inc, dk, ds in CLS_BLOCKS:
            s = "cls_scope/" + name
            h = T.residual_block(h, ps.w[s + "/Variable"], ps.w[s + "/Variable_1"], kp,
                                 ps.bn[s + "/" + name + "_1"], ps.bn[s + "/" + name + "_2"],
                                 inc_dim=inc, is_train=True, leak=True)
            h = T.conv_bn_relu2d(h, ps.w[s + "/Variable_2"], kp, ps.bn[s + "/" + name + "_3"],
                                 strides=(1, ds, ds, 1), is_train=True, leak=True)
        h = T.conv_bn_relu2d(h, ps.w["cls_scope/cls_6/Variable"], kp, ps.bn["cls_scope/cls_6/cls_6"],
                             padding="SYMMETRIC", strides=(1, 2, 2, 1), is_train=True, leak=True)
        flat = h.reshape(-1, 32 * FB * 4)
        return flat @ ps.w["cls_scope/cls_out/Variable"], d_input

    def mask_critic(self, logits):
        """adversarial.py:402-443."""
        ps, kp, m = self.ps, self.critic_keep_prob, "mask_cls_scope/"
        h = T.conv_bn_relu2d(logits, ps.w[m + "mask_cls_1/Variable"], kp, ps.bn[m + "mask_cls_1/mask_cls_1"],
                             strides=(1, 2, 2, 1), is_train=True, leak=True)
        h = T.residual_block(h, ps.w[m + "mask_cls_2/Variable"], ps.w[m + "mask_cls_2/Variable_1"], kp,
                             ps.bn[m + "mask_cls_2/m_cls_2_1"], ps.bn[m + "mask_cls_2/m_cls_2_2"], is_train=True, leak=True)
        h = T.conv_bn_relu2d(h, ps.w[m + "mask_cls_2/Variable_2"], kp, ps.bn[m + "mask_cls_2/m_cls_2_3"],
                             strides=(1, 4, 4, 1), is_train=True, leak=True)
        h = T.residual_block(h, ps.w[m + "mask_cls_3/Variable"], ps.w[m + "mask_cls_3/Variable_1"], kp,
                             ps.bn[m + "mask_cls_3/m_cls_3_1"], ps.bn[m + "mask_cls_3/m_cls_3_2"], inc_dim=True,
                             is_train=True, leak=True)
        h = T.conv_bn_relu2d(h, ps.w[m + "mask_cls_3/Variable_2"], kp, ps.bn[m + "mask_cls_3/m_cls_3_3"],
                             strides=(1, 4, 4, 1), is_train=True, leak=True)
        h = T.conv_bn_relu2d(h, ps.w[m + "mask_cls_4/Variable"], kp, ps.bn[m + "mask_cls_4/m_cls_4"],
                             padding="SYMMETRIC", strides=(1, 4, 4, 1), is_train=True, leak=True)
        return h.reshape(-1, 16 * FB * 4) @ ps.w[m + "m_cls_out/Variable"]

    def segment(self, x, stream, keep_prob, front_bn, joint_bn=False):
        """stream 'mr' -> group_1..6 / pred_* BN ; 'ct' -> adapt_1..6 / adapt_* BN; then shared back half."""
        if stream == "mr":
            c4, c6 = run_front(self.ps, x, "group_%d", self.pred_front, keep_prob, front_bn)
        else:
            c4, c6 = run_front(self.ps, x, "adapt_%d", self.adapt_front, keep_prob, front_bn)
        c9, b8, b7, logits = run_back(self.ps, c6, self.pred_back, keep_prob, joint_bn, self.batch_size, self.num_cls)
        return {"c4_2": c4, "c6_2": c6, "b7": b7, "b8": b8, "c9_2": c9, "logits": logits}

    def _l2(self, names):
        return sum(T.l2_loss(self.ps.w[n]) for n in names)

    # ---- losses (adversarial.py:445-476) --------------------------------------------------------
    def dis_losses(self, ct_cls, mr_cls, ct_m, mr_m):
        dis = -1 * self.miu_dis * (mr_cls - ct_cls).mean()
        m_dis = -1 * self.miu_dis * (mr_m - ct_m).mean() if ct_m is not None else 0.0
        # cls_weights / m_cls_weights are appended once per create_* call and each is called twice
        # (CT then MR, adversarial.py:98-99,118-119) => every weight is counted twice in the sums.
        dis_reg = self.gan_reg * self.miu_dis * 2 * self._l2(self.d_weights)
        m_reg = self.gan_reg * self.miu_dis * 2 * self._l2(self.m_weights)
        return dis + self.lam * m_dis, dis_reg + self.lam * m_reg

    def gen_losses(self, ct_cls, ct_m):
        gen = -1 * self.miu_gen * ct_cls.mean()
        m_gen = -1 * self.miu_gen * ct_m.mean() if ct_m is not None else 0.0
        gen_reg = self.gan_reg * self.miu_gen * self._l2(self.adapt_w)
        return gen + self.lam * m_gen, gen_reg

    # ---- steps ----------------------------------------------------------------------------------
    def d_step(self, mr, ct, keep_prob=0.75):
        """adversarial.py:840-862: all segmenter BN switches False, dropout on; minimise
        dis_loss + dis_reg/dis_sub_iter over cls_vars with RMSProp; then clip D/M conv+FC weights."""
        with torch.no_grad():
            fm = self.segment(mr, "mr", keep_prob, False)
            fc = self.segment(ct, "ct", keep_prob, False)
        ct_cls, _ = self.classifier(fc["c4_2"], fc["c6_2"], fc["b7"], fc["c9_2"], fc["logits"])
        mr_cls, _ = self.classifier(fm["c4_2"], fm["c6_2"], fm["b7"], fm["c9_2"], fm["logits"])
        if self.lam != 0:
            ct_m = self.mask_critic(fc["logits"])
            mr_m = self.mask_critic(fm["logits"])
        else:
            ct_m = mr_m = None
        dis_loss, dis_reg = self.dis_losses(ct_cls, mr_cls, ct_m, mr_m)
        total = dis_loss + dis_reg / self.dis_sub_iter
        grads = list(torch.autograd.grad(total, self.d_trainables, allow_unused=True))
        # lambda==0: TF still produces (zero) gradients for the mask critic through `0 * m_dis_loss`
        grads = [torch.zeros_like(p) if g is None else g for p, g in zip(self.d_trainables, grads)]
        self.d_opt.step(grads)
        with torch.no_grad():  # clip_op, adversarial.py:653-654: only names containing "Variable"
            for n in self.cls_w:
                self.ps.w[n].clamp_(-0.03, 0.03)
        return {"dis_loss": float(dis_loss), "dis_reg": float(dis_reg), "grads": grads,
                "ct_cls": ct_cls.detach(), "mr_cls": mr_cls.detach()}

    def g_step(self, ct, keep_prob=0.75):
        """adversarial.py:869-882: ct_front_bn True (DAM BN batch stats + moving update), others False;
        minimise ct_gen_loss + gen_reg/gen_sub_iter over adapt_vars."""
        fc = self.segment(ct, "ct", keep_prob, True)
        ct_cls, _ = self.classifier(fc["c4_2"], fc["c6_2"], fc["b7"], fc["c9_2"], fc["logits"])
        ct_m = self.mask_critic(fc["logits"]) if self.lam != 0 else None
        gen_loss, gen_reg = self.gen_losses(ct_cls, ct_m)
        total = gen_loss + gen_reg / self.gen_sub_iter
        grads = list(torch.autograd.grad(total, self.g_trainables, allow_unused=True))
        self.g_opt.step(grads)
        return {"gen_loss": float(gen_loss), "gen_reg": float(gen_reg), "grads": grads, "ct_cls": ct_cls.detach()}


# ---- synthetic inputs ---------------------------------------------------------------
def synthetic_images(B, seed, shift=0.0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 256, 256, 3, generator=g) * scale + shift


def synthetic_labels(B, seed, size=256, num_cls=5):
    """Deterministic nested-ellipse class maps: classes 1..4 nested on background 0."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.zeros((B, size, size), np.int64)
    for b in range(B):
        cy, cx = size / 2 + rng.uniform(-20, 20), size / 2 + rng.uniform(-20, 20)
        for c, rad in zip(range(1, num_cls), (100, 75, 50, 25)):
            ry, rx = rad * rng.uniform(0.8, 1.0), rad * rng.uniform(0.8, 1.0)
            out[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = c
    return out
