"""Topology of the PnP-AdaNet graphs, written as small token strings and executed through the
layers.py / ops.py surface (so variable names follow the reference's checkpoint naming contract).

tokens per group:  cN = conv2d 3x3 -> N (dropout only)      rN = residual_block N->N
                   RN = residual_block with inc_dim (N/2->N) dN = DR_block (rate 2) N->N
                   bN = conv_bn_relu2d (leaky) N->N          p  = max_pool2d 2

Segmenter front (groups 1-6, == the CT "DAM" adapt_1-6): source_segmenter.py:91-161, adversarial.py:130-269
Segmenter back  (groups 7-10 + output, shared):           source_segmenter.py:163-209, adversarial.py:273-318
"""
from . import layers as L
from . import ops
from . import runtime as rt

FB = 16
FRONT = {1: "c16 r16 p", 2: "R32 p", 3: "R64 r64 p", 4: "R128 r128", 5: "R256 r256", 6: "r256 r256"}
BACK = {7: "R512 r512", 8: "d512 d512", 9: "b512 b512"}


def _parse(tok):
    return tok[0], (int(tok[1:]) if len(tok) > 1 else 0)


class SegmenterHalf:
    """Creates the variables of a run of groups once (TF graph-build time) and replays them eagerly."""

    def __init__(self, groups, scope_fmt, in_channels, bn_namer, trainable, stddev, shared_from=None):
        """scope_fmt: 'group_%d' | 'adapt_%d'.  bn_namer(group, block_idx, kind) -> BN scope base name (None =
        anonymous 'BatchNorm_k').  shared_from: groups >= this index use tf.get_variable naming (same result)."""
        self.groups = groups
        self.scope_fmt = scope_fmt
        self.plan = []          # (group, scope, [(kind, weights..., bn scope(s), cin, cout)])
        self.weights = []
        c = in_channels
        for gi in sorted(groups):
            scope = scope_fmt % gi
            ops_ = []
            blk = 0
            j = 0
            with rt.variable_scope(scope):
                for tok in groups[gi].split():
                    kind, n = _parse(tok)
                    if kind == "p":
                        ops_.append(("p",))
                        continue
                    blk += 1

                    def var(shape, j=j):
                        name = "Variable" if j == 0 else "Variable_%d" % j
                        return L.sharable_weight_variable(shape, stddev=stddev, trainable=trainable, name=name)
                    if kind == "c":
                        w = var([3, 3, c, n])
                        j += 1
                        ops_.append(("c", w))
                        self.weights.append(w)
                        blk -= 1
                    elif kind == "b":
                        w = var([3, 3, c, n])
                        j += 1
                        s = bn_namer(gi, blk, "b")
                        L.bn_variables(s, n, trainable)
                        ops_.append(("b", w, s))
                        self.weights.append(w)
                    else:
                        w1 = var([3, 3, c, n])
                        j += 1
                        w2 = var([3, 3, n, n], j)
                        j += 1
                        s = bn_namer(gi, blk, kind)
                        L.bn_variables(s[0], n, trainable)
                        L.bn_variables(s[1], n, trainable)
                        ops_.append((kind, w1, w2, s))
                        self.weights += [w1, w2]
                    c = n
            self.plan.append((gi, scope, ops_))
        self.out_channels = c

    def run(self, x, keep_prob, is_train, bn_trainable=True):
        """returns (output, {group: activation})"""
        taps = {}
        h = x
        for gi, scope, ops_ in self.plan:
            with rt.variable_scope(scope):
                for op in ops_:
                    k = op[0]
                    if k == "p":
                        h = L.max_pool2d(h, 2)
                    elif k == "c":
                        h = L.conv2d(h, op[1], keep_prob)
                    elif k == "b":
                        h = L.conv_bn_relu2d(h, op[1], keep_prob, is_train=is_train, scope=op[2], bn_trainable=bn_trainable, leak=True)
                    else:
                        cfg = dict(keep_prob=keep_prob, is_train=is_train, bn_trainable=bn_trainable, leak=True)
                        h = _two_conv_block(h, op, cfg)
            taps[gi] = h
        return h, taps


def _two_conv_block(h, op, cfg):
    """residual_block / DR_block with explicit per-conv BN scopes (the reference derives them as
    scope+'_1' / scope+'_2' (layers.py:151-155); anonymous scopes are 'BatchNorm_k' pairs)."""
    from . import functional as F
    kind, w1, w2, (s1, s2) = op
    inc = kind == "R"
    dil = 2 if kind == "d" else 1
    cin = h.shape[-1]
    bn1 = L.bn_variables(s1, w1.shape[3], cfg["bn_trainable"])
    bn2 = L.bn_variables(s2, w2.shape[3], cfg["bn_trainable"])
    c1 = F.LayerCfg(dil=dil, keep_prob=cfg["keep_prob"], bn=bn1, bn_training=cfg["is_train"], act=F.ACT_LRELU)
    c2 = F.LayerCfg(dil=dil, keep_prob=cfg["keep_prob"], bn=bn2, bn_training=cfg["is_train"], act=F.ACT_LRELU,
                    skip_off=(cin // 2 if inc else 0))
    return F.res_block(h, w1, w2, c1, c2)


class SegmenterTail:
    """group_10 (3x3 SYMMETRIC 512 -> 64*8*num_cls, dropout) -> PS r=8 -> output (5x5 SYMMETRIC -> num_cls,
    keep_prob 1).  source_segmenter.py:195-207, adversarial.py:306-316."""

    def __init__(self, num_cls, trainable, stddev):
        self.num_cls = num_cls
        with rt.variable_scope("group_10"):
            self.w10 = L.sharable_weight_variable([3, 3, FB * 32, 64 * num_cls * 8], stddev=stddev, trainable=trainable, name="Variable")
        with rt.variable_scope("output"):
            self.w11 = L.sharable_weight_variable([5, 5, num_cls * 8, num_cls], stddev=stddev, trainable=trainable, name="Variable")
        self.weights = [self.w10, self.w11]

    def run(self, c9, keep_prob, batch_size):
        from . import functional as F
        conv10 = L.conv2d(c9, self.w10, keep_prob_=keep_prob, padding='SYMMETRIC')
        import torch
        if F.FUSE_TAIL and not (self.w11.requires_grad and torch.is_grad_enabled()) and self.w11.shape[3] in (5, 8):
            # output filter takes no gradient (adversarial steps, any torch.no_grad() forward): PS + mirror pad + convolution in one kernel
            ops._check(conv10, batch_size)
            return F.tail_ps_conv(conv10, self.w11, 8, self.num_cls * 8, batch_size)
        flat = ops.PS(conv10, r=8, n_channel=self.num_cls * 8, batch_size=batch_size)
        return L.conv2d(flat, self.w11, keep_prob_=1., padding='SYMMETRIC')
