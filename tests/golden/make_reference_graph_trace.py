"""Trace the REFERENCE'S OWN graph-construction code (adversarial.py:44-443 + layers.py + ops.py, imported unmodified from
/root/reference) under a recording shim of `tensorflow`, and commit the canonical layer list as
tests/golden/reference_graph_trace.json.

Nothing numerical runs: the shim's tensors carry only a static shape, a provenance tag and the variable they came from.  What
the trace pins is the ARCHITECTURE the reference builds -- for every convolution its filter variable (name and shape), stride,
dilation, padding rule, where dropout sits and which keep_prob feeds it, the batch-norm scope / training switch / trainable flag,
the skip connection (identity or zero channel pad), the activation; the pooling positions; every PS call; the channel layout of
the discriminator input; the final matmuls -- in execution order, for the zip network (MR path + CT "DAM" path), both calls of
create_second_half, both calls of create_classifier and both calls of create_mask_critic.

`Full_DRN.__init__` of the reference raises AttributeError at adversarial.py:102 (`self.predicter`) after the segmenter halves
and the feature discriminator have been built; the mask critic is then traced by calling `create_mask_critic` exactly as
adversarial.py:116-118 would.  tests/test_reference_graph_trace.py replays the PRODUCT's graph code on CPU with recording stand-ins
for its kernels and compares the two lists.

    python tests/golden/make_reference_graph_trace.py          # needs /root/reference; the tests only read the .json
"""
import contextlib
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_graph_trace.json")
BATCH = 2          # any B >= 2 (the B == 1 branch of ops.py is a different permutation, not a different architecture)


# ------------------------------------------------------------------------------------------------
# symbolic tensors
# ------------------------------------------------------------------------------------------------
class _Shape(object):
    def __init__(self, s):
        self._s = list(s)

    def as_list(self):
        return list(self._s)


class Sym(object):
    """static shape + provenance.  `src` names what produced it (used to label the discriminator input channels)."""
    _n = 0

    def __init__(self, shape, src, var=None):
        self.shape = [int(v) for v in shape]
        self.src = src
        self.var = var              # Variable record if this tensor IS a variable
        Sym._n += 1
        self.id = Sym._n

    def get_shape(self):
        return _Shape(self.shape)

    def _bin(self, other, op):
        o = other.src if isinstance(other, Sym) else repr(other)
        out = Sym(self.shape, "%s(%s,%s)" % (op, self.src, o))
        REC.raw(op, a=self, b=other, out=out)
        return out

    def __add__(self, o):
        return self._bin(o, "add")

    __radd__ = __add__

    def __mul__(self, o):
        return self._bin(o, "mul")

    __rmul__ = __mul__

    def __sub__(self, o):
        return self._bin(o, "sub")

    def __truediv__(self, o):
        return self._bin(o, "div")

    __div__ = __truediv__

    def __neg__(self):
        return self._bin(-1.0, "mul")

    def __rsub__(self, o):
        return self._bin(o, "sub")

    def __rtruediv__(self, o):
        return self._bin(o, "div")

    __rdiv__ = __rtruediv__

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        shape = []
        for n, i in zip(self.shape, idx):
            if isinstance(i, slice):
                shape.append(len(range(*i.indices(n))))
        shape += self.shape[len(idx):]
        return Sym(shape, self.src)


class Flag(object):
    """a scalar placeholder (BN training switch, keep_prob)"""

    def __init__(self, name, default=None):
        self.name, self.default = name, default

    def tag(self):
        return "ph:" + self.name


def _short(src):
    """compact provenance: the variable of the last convolution that produced the tensor"""
    if src.startswith("PS("):
        return "PS(%s)" % _short(src[3:-1])
    if src.startswith("argmax("):
        return "argmax(%s)" % _short(src[7:-1])
    i = max(src.rfind("conv("), src.rfind("aconv("))
    if i < 0:
        return src
    j = src.index("(", i)
    return "out_of:" + src[j + 1:src.index(")", j)]


def _tag(v):
    if isinstance(v, Flag):
        return v.tag()
    if isinstance(v, bool):
        return v
    if isinstance(v, (int, float)):
        return float(v)
    return repr(v)


# ------------------------------------------------------------------------------------------------
# recorder: raw op stream -> canonical layer events
# ------------------------------------------------------------------------------------------------
class Recorder(object):
    def __init__(self):
        self.events = []          # canonical
        self.section = None
        self.depth_ps = 0
        self.open = None          # conv event still collecting dropout / bn / skip / act
        self.pending_pad = None   # SYMMETRIC tf.pad feeding the next conv
        self.pending_skip = {}    # Sym id of a zero-channel-padded skip -> pad amount

    def set_section(self, name):
        self._close()
        self.section = name

    def emit(self, ev):
        self._close()
        ev["section"] = self.section
        self.events.append(ev)

    def _close(self):
        if self.open is not None:
            ev, self.open = self.open, None
            ev.pop("_cur", None)
            ev["section"] = self.section
            self.events.append(ev)

    def raw(self, op, **kw):
        if self.depth_ps:
            return
        if op == "pad":
            x, paddings, mode, out = kw["x"], kw["paddings"], kw["mode"], kw["out"]
            if mode == "SYMMETRIC":
                self.pending_pad = (out.id, paddings)
            else:
                assert paddings[0] == [0, 0] and paddings[1] == [0, 0] and paddings[2] == [0, 0], paddings
                assert paddings[3][0] == paddings[3][1]
                self.pending_skip[out.id] = (x.id, paddings[3][0])
            return
        if op in ("conv2d", "atrous_conv2d"):
            self._close()
            x, W = kw["x"], kw["W"]
            padding = kw["padding"]
            in_shape = x.shape
            if self.pending_pad is not None and self.pending_pad[0] == x.id:
                assert padding == "VALID"
                p = self.pending_pad[1]
                assert p[1][0] == p[1][1] == W.shape[0] // 2 and p[2][0] == p[2][1] == W.shape[1] // 2, (p, W.shape)
                padding = "SYMMETRIC"
                in_shape = [x.shape[0], x.shape[1] - 2 * p[1][0], x.shape[2] - 2 * p[2][0], x.shape[3]]
            self.pending_pad = None
            self.open = {"op": "conv", "w": W.var["name"], "wshape": list(W.shape), "w_trainable": W.var["trainable"],
                         "w_kind": W.var["kind"], "w_stddev": W.var["stddev"],
                         "stride": kw["stride"], "dil": kw["rate"], "padding": padding, "in": in_shape[1:], "out": kw["out"].shape[1:],
                         "input_src": x.src if x.src.startswith("ph:") else None,
                         "input_layout": ([[_short(s_), int(c_)] for s_, c_ in x.parts] if len(getattr(x, "parts", [])) > 1 else None),
                         "keep": None, "bn": None, "bn_train": None, "bn_trainable": None, "bn_decay": None, "skip": "none", "act": "none",
                         "_cur": kw["out"].id, "_x": x.id}
            return
        ev = self.open
        if op == "dropout":
            assert ev is not None and kw["x"].id == ev["_cur"], "dropout not directly after a convolution"
            assert ev["bn"] is None and ev["keep"] is None
            ev["keep"] = _tag(kw["keep_prob"])
            ev["_cur"] = kw["out"].id
            return
        if op == "batch_norm":
            assert ev is not None and kw["x"].id == ev["_cur"], "batch_norm not on the conv/dropout output"
            ev["bn"], ev["bn_train"], ev["bn_trainable"], ev["bn_decay"] = kw["scope"], _tag(kw["is_training"]), kw["trainable"], kw["decay"]
            ev["_cur"] = kw["out"].id
            return
        if op == "add":
            a, b = kw["a"], kw["b"]
            if not (ev is not None and isinstance(b, Sym) and b.id == ev["_cur"] and isinstance(a, Sym) and a.shape == b.shape
                    and ev["bn"] is not None):
                return            # loss arithmetic, not the `x_s + _inner_conv` of layers.py:164-166 / 186-189
            if a.id in self.pending_skip:
                ev["skip"] = "pad%d" % self.pending_skip[a.id][1]
            else:
                ev["skip"] = "identity"
            ev["skip_from_conv_input_of"] = None
            ev["_cur"] = kw["out"].id
            return
        if op in ("relu", "leaky_relu"):
            assert ev is not None and kw["x"].id == ev["_cur"], "activation not on the layer output"
            ev["act"] = "relu" if op == "relu" else "lrelu%g" % kw["alpha"]
            ev["_cur"] = kw["out"].id
            self._close()
            return
        if op == "max_pool":
            self.emit({"op": "maxpool", "k": kw["k"], "stride": kw["stride"], "padding": kw["padding"], "in": kw["x"].shape[1:],
                       "out": kw["out"].shape[1:]})
            return
        if op == "matmul":
            self.emit({"op": "fc", "w": kw["W"].var["name"], "wshape": list(kw["W"].shape), "w_trainable": kw["W"].var["trainable"],
                       "in": kw["x"].shape[1:]})
            return
        if op in ("tile", "concat", "argmax", "expand_dims", "cast", "reshape", "exp", "reduce_sum", "div", "clip", "stack", "mul", "sub", "add"):
            return            # bookkeeping handled through Sym.src (discriminator input layout) or irrelevant here
        raise RuntimeError("unhandled raw op %s" % op)


REC = Recorder()


# ------------------------------------------------------------------------------------------------
# the tensorflow shim
# ------------------------------------------------------------------------------------------------
class Graph(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.vars = {}                # full name -> record
        self.order = []
        self.scope = []               # variable-scope stack: names of tf.get_variable / batch_norm variables
        self.nscope = []              # name-scope stack (variable scopes open one too): names of tf.Variable ops
        self.used = {}                # ("v"|"n", scope path) -> {base name: count}   (TF's unique-name rule)

    def path(self, kind="v"):
        return "/".join(s for s in (self.scope if kind == "v" else self.nscope) if s)

    def unique(self, base, kind="v"):
        d = self.used.setdefault((kind, self.path(kind)), {})
        n = d.get(base, 0)
        d[base] = n + 1
        return base if n == 0 else "%s_%d" % (base, n)

    def full(self, name, kind="v"):
        p = self.path(kind)
        return (p + "/" + name) if p else name

    def new_var(self, full, shape, trainable, kind, stddev=None):
        rec = {"name": full, "shape": [int(v) for v in shape], "trainable": bool(trainable), "kind": kind, "stddev": stddev}
        assert full not in self.vars, full
        self.vars[full] = rec
        self.order.append(full)
        return rec


G = Graph()


class _Init(object):
    def __init__(self, shape, stddev):
        self.shape, self.stddev = shape, stddev


def _conv_out(n, k, s, d, padding):
    if padding == "SAME":
        return -(-n // s)
    return (n - ((k - 1) * d + 1)) // s + 1


def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.int64, tf.string, tf.AUTO_REUSE = "float32", "int32", "int64", "string", "AUTO_REUSE"
    tf.FixedLenFeature = lambda shape, dtype: ("FixedLenFeature", shape, dtype)
    tf.reset_default_graph = G.reset

    def placeholder(dtype, shape=None, name=None):
        if shape is None:
            ph = Flag(name or "keep_prob")
            return ph
        placeholder.count += 1
        nm = name or "Placeholder_%d" % placeholder.count
        return Sym([BATCH if v is None else v for v in shape], "ph:" + nm)
    placeholder.count = 0
    tf.placeholder = placeholder
    tf.placeholder_with_default = lambda default, shape=None, name=None: Flag(name, default)

    @contextlib.contextmanager
    def variable_scope(name, reuse=None):
        G.scope.append(name)
        G.nscope.append(name)
        try:
            yield name
        finally:
            G.scope.pop()
            G.nscope.pop()
    tf.variable_scope = variable_scope

    @contextlib.contextmanager
    def name_scope(name):
        G.nscope.append(name)          # affects tf.Variable names only; tf.get_variable (batch_norm) ignores it
        try:
            yield name
        finally:
            G.nscope.pop()
    tf.name_scope = name_scope

    tf.truncated_normal = lambda shape, stddev=1.0: _Init(shape, stddev)
    tf.truncated_normal_initializer = lambda stddev=1.0: _Init(None, stddev)

    def Variable(initial, trainable=True, name=None):
        assert isinstance(initial, _Init)
        full = G.full(G.unique(name or "Variable", "n"), "n")
        rec = G.new_var(full, initial.shape, trainable, "tf.Variable", initial.stddev)
        return Sym(initial.shape, "var:" + full, var=rec)
    tf.Variable = Variable

    def get_variable(name, shape=None, initializer=None, trainable=True):
        full = G.full(name)
        if full in G.vars:                   # AUTO_REUSE
            rec = G.vars[full]
            assert rec["shape"] == [int(v) for v in shape], (full, rec["shape"], shape)
        else:
            rec = G.new_var(full, shape, trainable, "tf.get_variable", initializer.stddev if initializer else None)
        return Sym(shape, "var:" + full, var=rec)
    tf.get_variable = get_variable

    tf.constant = lambda value, shape=None: value
    tf.cast = lambda x, dtype: (REC.raw("cast", x=x) or x) if isinstance(x, Sym) else x

    def pad(x, paddings, mode="CONSTANT"):
        p = [[int(a), int(b)] for a, b in paddings]
        out = Sym([n + a + b for n, (a, b) in zip(x.shape, p)], x.src if mode != "SYMMETRIC" else "mirror(%s)" % x.src)
        REC.raw("pad", x=x, paddings=p, mode=mode, out=out)
        return out
    tf.pad = pad

    nn = types.ModuleType("tensorflow.nn")

    def conv2d(x, W, strides, padding):
        assert strides[0] == strides[3] == 1 and strides[1] == strides[2]
        s = int(strides[1])
        kh, kw, ci, co = W.shape
        assert ci == x.shape[3], (W.var["name"], x.shape, W.shape)
        out = Sym([x.shape[0], _conv_out(x.shape[1], kh, s, 1, padding), _conv_out(x.shape[2], kw, s, 1, padding), co], "conv(%s)" % W.var["name"])
        REC.raw("conv2d", x=x, W=W, stride=s, rate=1, padding=padding, out=out)
        return out
    nn.conv2d = conv2d

    def atrous_conv2d(x, W, rate, padding):
        kh, kw, ci, co = W.shape
        assert ci == x.shape[3]
        out = Sym([x.shape[0], _conv_out(x.shape[1], kh, 1, rate, padding), _conv_out(x.shape[2], kw, 1, rate, padding), co], "aconv(%s)" % W.var["name"])
        REC.raw("atrous_conv2d", x=x, W=W, stride=1, rate=int(rate), padding=padding, out=out)
        return out
    nn.atrous_conv2d = atrous_conv2d

    def dropout(x, keep_prob):
        out = Sym(x.shape, x.src)
        REC.raw("dropout", x=x, keep_prob=keep_prob, out=out)
        return out
    nn.dropout = dropout

    def relu(x):
        out = Sym(x.shape, x.src)
        REC.raw("relu", x=x, out=out)
        return out
    nn.relu = relu

    def leaky_relu(x, alpha=0.2):
        out = Sym(x.shape, x.src)
        REC.raw("leaky_relu", x=x, alpha=alpha, out=out)
        return out
    nn.leaky_relu = leaky_relu

    def max_pool(x, ksize, strides, padding):
        k, s = int(ksize[1]), int(strides[1])
        out = Sym([x.shape[0], -(-x.shape[1] // s), -(-x.shape[2] // s), x.shape[3]], "pool(%s)" % x.src)
        REC.raw("max_pool", x=x, k=k, stride=s, padding=padding, out=out)
        return out
    nn.max_pool = max_pool
    nn.l2_loss = lambda w: Sym([], "l2")
    nn.softmax = lambda x: Sym(x.shape, x.src)
    tf.nn = nn

    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")

    def batch_norm(x, is_training=True, decay=0.999, scale=False, center=True, scope=None, variables_collections=None,
                   updates_collections="UPDATE_OPS", trainable=True):
        assert scale and center and updates_collections is None
        sc = scope if scope is not None else G.unique("BatchNorm")
        G.scope.append(sc)
        try:
            C = x.shape[-1]
            for nm, tr in (("beta", trainable), ("gamma", trainable), ("moving_mean", False), ("moving_variance", False)):
                full = G.full(nm)
                if full not in G.vars:
                    G.new_var(full, [C], tr, "batch_norm")
                else:
                    assert G.vars[full]["shape"] == [C]
            full_scope = G.path()
        finally:
            G.scope.pop()
        out = Sym(x.shape, x.src)
        REC.raw("batch_norm", x=x, scope=full_scope, is_training=is_training, trainable=bool(trainable), decay=decay, out=out)
        return out
    layers.batch_norm = batch_norm
    contrib.layers = layers
    framework = types.ModuleType("tensorflow.contrib.framework")
    contrib.framework = framework
    tf.contrib = contrib

    # ---- shape ops (PS, discriminator input assembly, softmax bookkeeping) ----
    def reshape(x, shape):
        shape = list(shape)
        n = int(np.prod(x.shape))
        if -1 in shape:
            known = int(np.prod([v for v in shape if v != -1]))
            shape[shape.index(-1)] = n // known
        assert int(np.prod(shape)) == n, (x.shape, shape)
        out = Sym(shape, x.src)
        REC.raw("reshape", x=x, out=out)
        return out
    tf.reshape = reshape
    tf.transpose = lambda x, perm: Sym([x.shape[p] for p in perm], x.src)

    def split(value, num, axis):
        axis = axis % len(value.shape)
        assert value.shape[axis] % num == 0
        s = list(value.shape)
        s[axis] //= num
        return [Sym(s, value.src) for _ in range(num)]
    tf.split = split

    def concat(values, axis, name=None):
        axis = axis % len(values[0].shape)
        s = list(values[0].shape)
        s[axis] = sum(v.shape[axis] for v in values)
        srcs = []
        for v in values:
            srcs += getattr(v, "parts", [(v.src, v.shape[axis])])
        out = Sym(s, "concat")
        out.parts = srcs if axis == len(s) - 1 else [(values[0].src, s[-1])]
        if axis != len(s) - 1:
            out.src = values[0].src
        REC.raw("concat", out=out)
        return out
    tf.concat = concat
    tf.squeeze = lambda x: Sym([v for v in x.shape if v != 1], x.src)

    def expand_dims(x, axis):
        s = list(x.shape)
        s.insert(axis if axis >= 0 else len(s) + 1 + axis, 1)
        out = Sym(s, x.src)
        return out
    tf.expand_dims = expand_dims

    def tile(x, multiples):
        if isinstance(multiples, Sym):
            return Sym(x.shape[:3] + [x.shape[3] * 0 + 1 * x.shape[3]], x.src)       # pixel_wise_softmax_2 bookkeeping only
        out = Sym([n * int(m) for n, m in zip(x.shape, multiples)], x.src)
        out.parts = [(x.src, x.shape[-1])] * int(multiples[-1])
        return out
    tf.tile = tile

    def argmax(x, axis):
        s = list(x.shape)
        s.pop(axis)
        return Sym(s, "argmax(%s)" % x.src)
    tf.argmax = argmax
    tf.shape = lambda x: list(x.shape)
    tf.equal = lambda a, b: True
    tf.stack = lambda vals: Sym([len(vals)], "stack")
    tf.exp = lambda x: Sym(x.shape, x.src)
    tf.log = lambda x: Sym(x.shape, x.src)

    def reduce_sum(x, axis=None, keep_dims=False):
        if axis is None:
            return Sym([], "sum")
        s = list(x.shape)
        if keep_dims:
            s[axis] = 1
        else:
            s.pop(axis)
        return Sym(s, x.src)
    tf.reduce_sum = reduce_sum
    tf.reduce_mean = lambda x, axis=None, name=None: Sym([], "mean")
    tf.div = lambda a, b, name=None: Sym(a.shape, a.src)
    tf.add = lambda a, b: Sym(a.shape, a.src)
    tf.clip_by_value = lambda x, lo, hi, name=None: Sym(x.shape, x.src)
    tf.reverse = lambda x, dims: Sym(x.shape, x.src)
    tf.slice = lambda x, begin, size: x

    def matmul(x, W):
        assert x.shape[-1] == W.shape[0]
        out = Sym([x.shape[0], W.shape[1]], "fc(%s)" % W.var["name"])
        REC.raw("matmul", x=x, W=W, out=out)
        return out
    tf.matmul = matmul
    tf.one_hot = lambda x, depth, axis=-1: Sym(list(x.shape) + [depth], x.src)
    tf.confusion_matrix = lambda a, b, num_classes: Sym([num_classes, num_classes], "cm")

    python = types.ModuleType("tensorflow.python")
    python.debug = types.ModuleType("tensorflow.python.debug")
    tf.python = python
    return tf, {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers,
                "tensorflow.python": python, "tensorflow.python.debug": python.debug}


def _load(name, alias=None):
    spec = importlib.util.spec_from_file_location(alias or name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias or name] = mod
    spec.loader.exec_module(mod)
    return mod


def _layout(sym):
    """channel layout of the discriminator input: [(source tag, channels), ...] with consecutive equal sources merged"""
    out = []
    for src, c in getattr(sym, "parts", [(sym.src, sym.shape[-1])]):
        out.append([src, c])
    return out


def main():
    tf, mods = _make_tf()
    sys.modules.update(mods)
    for name in ("nibabel", "matplotlib"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pnp_trace_")
    os.chdir(tmp)                     # adversarial.py opens a log file in the cwd at import time
    try:
        layers = _load("layers")
        ops = _load("ops")

        real_ps = ops.PS

        def PS(X, r, n_channel=8, batch_size=10):
            REC._close()
            REC.depth_ps += 1
            try:
                out = real_ps(X, r, n_channel=n_channel, batch_size=batch_size)
            finally:
                REC.depth_ps -= 1
            assert out.shape == [X.shape[0], X.shape[1] * r, X.shape[2] * r, n_channel], (X.shape, out.shape)
            out.src = "PS(%s)" % X.src
            out.parts = [(out.src, n_channel)]
            REC.emit({"op": "PS", "r": int(r), "n_channel": int(n_channel), "batch_size_arg": "self.batch_size" if batch_size == BATCH else int(batch_size),
                      "in": X.shape[1:], "out": out.shape[1:], "of": _short(X.src)})
            return out
        ops.PS = PS
        _load("lib")
        adv = _load("adversarial")

        cfg = {"mr_front_trainable": False, "ct_front_trainable": True, "joint_trainable": False, "cls_trainable": True, "m_cls_trainable": True,
               "restore_skip_kwd": ["Adam", "RMS", "cls"]}                     # train_gan.py:39-46 (train-gan phase)
        cost_kwargs = {"regularizer": 1e-4, "gan_regularizer": 1e-4, "miu_dis": 1e-3, "miu_gen": 2e-3, "lambda_mask_loss": 0.1}

        # mark sections by wrapping the four graph-construction methods
        calls = {}

        def sectioned(fn, label):
            def wrapper(self, *a, **k):
                calls[label] = calls.get(label, 0) + 1
                REC.set_section("%s#%d" % (label, calls[label]))
                out = fn(self, *a, **k)
                REC._close()
                if label == "create_classifier":
                    pass
                return out
            return wrapper
        for label in ("create_zip_network", "create_second_half", "create_classifier", "create_mask_critic"):
            setattr(adv.Full_DRN, label, sectioned(getattr(adv.Full_DRN, label), label))

        # the discriminator input: capture the tensor entering cls_1 (first conv of each create_classifier call)
        net = adv.Full_DRN.__new__(adv.Full_DRN)
        err = None
        try:
            adv.Full_DRN.__init__(net, channels=3, n_class=5, batch_size=BATCH, cost_kwargs=cost_kwargs, network_config=cfg)
        except AttributeError as e:           # adversarial.py:102  self.predicter
            err = str(e)
        assert err is not None and "predicter" in err, err
        REC._close()
        # adversarial.py:116-118, executed the way __init__ would have
        with tf.variable_scope("mask_cls_scope", reuse=tf.AUTO_REUSE):
            net.create_mask_critic(Sym([BATCH, 256, 256, 5], "ct_logits"), num_cls=5)
            net.create_mask_critic(Sym([BATCH, 256, 256, 5], "mr_logits"), num_cls=5)
        REC._close()
    finally:
        os.chdir(cwd)

    for ev in REC.events:
        ev.pop("_x", None)
        ev.pop("skip_from_conv_input_of", None)
    trace = {"batch": BATCH,
             "init_error_after_classifier": err,
             "events": REC.events,
             "variables": [G.vars[n] for n in G.order],
             "weight_lists": {k: [s.var["name"] for s in getattr(net, k)] for k in
                              ("mr_front_weights", "ct_front_weights", "cls_weights", "m_cls_weights", "joint_weights")}}
    trace["var_groups"], trace["optimizer"], trace["cost_numeric"] = trace_training_wiring(tf, adv, net, trace["weight_lists"])
    trace["schedule"] = trace_training_schedule(tf, adv, net)
    trace["input_pipeline"] = trace_input_pipeline(tf, adv)
    trace["source_segmenter"] = trace_source_segmenter(tf)
    G_adv_order = [v["name"] for v in trace["variables"]]
    trace["transplant"] = trace_transplant(tf, adv, net, G_adv_order, [v["name"] for v in trace["source_segmenter"]["variables"]],
                                           {v["name"]: v["trainable"] for v in trace["variables"]})
    with open(OUT, "w") as f:
        json.dump(trace, f, indent=0, sort_keys=True)
    print("wrote %s: %d + %d events, %d + %d variables" % (OUT, len(trace["events"]), len(trace["source_segmenter"]["events"]),
                                                           len(trace["variables"]), len(trace["source_segmenter"]["variables"])))


class _Num(object):
    """temporarily turn the shim numeric: the reference's loss code is plain arithmetic over a handful of reductions"""

    def __init__(self, tf):
        self.tf = tf
        self.saved = {}

    def __enter__(self):
        tf = self.tf
        num = {"reduce_mean": lambda x, axis=None, name=None: np.mean(x), "reduce_sum": lambda x, axis=None, keep_dims=False: np.sum(x),
               "log": np.log, "clip_by_value": lambda x, lo, hi, name=None: np.clip(x, lo, hi),
               "Variable": lambda initial, trainable=True, name=None: float(initial)}
        for k, v in num.items():
            self.saved[k] = getattr(tf, k)
            setattr(tf, k, v)
        self.saved_nn = (tf.nn.l2_loss, tf.nn.softmax)
        tf.nn.l2_loss = lambda w: float(w.l2) if hasattr(w, "l2") else float(np.sum(np.asarray(w, np.float64) ** 2) / 2)

        def softmax(x):
            e = np.exp(x - x.max(axis=-1, keepdims=True))
            return e / e.sum(axis=-1, keepdims=True)
        tf.nn.softmax = softmax
        return self

    def __exit__(self, *a):
        for k, v in self.saved.items():
            setattr(self.tf, k, v)
        self.tf.nn.l2_loss, self.tf.nn.softmax = self.saved_nn


class _VarObj(object):
    def __init__(self, rec, l2=None):
        self.name = rec["name"] + ":0"
        self.rec = rec
        self.l2 = l2


def trace_training_wiring(tf, adv, net, weight_lists):
    """adversarial.py:445-501 (cost, variable groups) and :633-656 (optimizers, clip) executed on the traced graph"""
    # ---- _get_variables_by_scope: membership by substring, in tf.global_variables() order
    tf.contrib.framework.get_variables = lambda: [_VarObj(G.vars[n]) for n in G.order]
    adv.Full_DRN._get_variables_by_scope(net)
    groups = {k: [v.rec["name"] for v in getattr(net, k)] for k in ("adapt_vars", "cls_vars", "seg_vars", "mri_seg_vars")}

    # ---- _get_optimizer with recording optimizers
    opt_log = {"optimizers": [], "clip": []}

    class _Loss(object):
        """linear combination of named scalars"""

        def __init__(self, terms):
            self.terms = dict(terms)

        def __add__(self, o):
            t = dict(self.terms)
            for k, v in (o.terms if isinstance(o, _Loss) else {"const": o}).items():
                t[k] = t.get(k, 0.0) + v
            return _Loss(t)

        __radd__ = __add__

        def __mul__(self, c):
            return _Loss({k: v * float(c) for k, v in self.terms.items()})

        __rmul__ = __mul__

    class RMSPropOptimizer(object):
        def __init__(self, learning_rate=None, **kw):
            self.cfg = {"kind": "RMSPropOptimizer", "learning_rate": learning_rate, "kwargs": dict(kw)}

        def minimize(self, loss, global_step=None, var_list=None):
            rec = dict(self.cfg)
            rec["objective"] = loss.terms
            rec["var_list"] = [v.rec["name"] for v in var_list]
            opt_log["optimizers"].append(rec)
            return "train_op_%d" % len(opt_log["optimizers"])
    train = types.ModuleType("tensorflow.train")
    train.RMSPropOptimizer = RMSPropOptimizer
    tf.train = train
    saved_var, saved_assign, saved_clip = tf.Variable, getattr(tf, "assign", None), tf.clip_by_value
    tf.Variable = lambda initial, trainable=True, name=None: float(initial)
    tf.clip_by_value = lambda v, lo, hi, name=None: ("clip", v, float(lo), float(hi))
    tf.assign = lambda var, val: opt_log["clip"].append({"var": var.rec["name"], "lo": val[2], "hi": val[3], "of_same_var": val[1] is var})
    net.dis_loss, net.dis_reg = _Loss({"dis_loss": 1.0}), _Loss({"dis_reg": 1.0})
    net.ct_gen_loss, net.gen_reg = _Loss({"ct_gen_loss": 1.0}), _Loss({"gen_reg": 1.0})
    me = types.SimpleNamespace(opt_kwargs={"learning_rate": 3e-4}, net=net, train_config={"dis_sub_iter": 20, "gen_sub_iter": 1})
    adv.Trainer._get_optimizer(me, 200, "global_step")
    opt_log["learning_rate_node"] = me.learning_rate_node
    opt_log["train_config_used"] = {"dis_sub_iter": 20, "gen_sub_iter": 1}
    tf.Variable, tf.clip_by_value = saved_var, saved_clip
    if saved_assign is not None:
        tf.assign = saved_assign

    # ---- _get_cost evaluated numerically: critic outputs and per-variable l2 values are the inputs
    rng = np.random.RandomState(7)
    uniq = sorted({n for lst in weight_lists.values() for n in lst})
    l2 = {n: float(rng.uniform(0.5, 20.0)) for n in uniq}
    for key, lst in weight_lists.items():
        setattr(net, key, [_VarObj({"name": n}, l2[n]) for n in lst])          # same multiplicities the graph code produced
    cls = {k: rng.standard_normal((BATCH, 1)) for k in ("ct_cls", "mr_cls", "ct_mask", "mr_mask")}
    cases = []
    for lam in (0.3, 0.0, None):
        ck = {"regularizer": 1e-4, "gan_regularizer": 1e-4, "miu_dis": 0.002, "miu_gen": 0.002}
        if lam is not None:
            ck["lambda_mask_loss"] = lam
        with _Num(tf):
            out = adv.Full_DRN._get_cost(net, None, None, cls["ct_cls"], cls["mr_cls"], cls["ct_mask"], cls["mr_mask"], dict(ck))
        cases.append({"cost_kwargs": ck, "dis_loss": float(out[0]), "gen_loss": float(out[1]), "fixed_coeff_reg": float(out[2]),
                      "dis_reg": float(out[3]), "gen_reg": float(out[4])})
    cost = {"l2": l2, "critic_outputs": {k: v.reshape(-1).tolist() for k, v in cls.items()}, "cases": cases}
    return groups, opt_log, cost


def trace_training_schedule(tf, adv, net):
    """Trainer.train (adversarial.py:767-946) inherited VERBATIM by a subclass that only replaces the data / monitoring
    helpers (next_batch, output_minibatch_stats, _initialize's summaries) and runs against a recording Session: what comes out
    is the order of optimizer / clip runs and the feed_dict of each (which BN switches and keep_prob every step kind uses)."""
    log = []
    names = {id(net.mr): "mr", id(net.ct): "ct", id(net.mr_front_bn): "mr_front_bn", id(net.joint_bn): "joint_bn",
             id(net.ct_front_bn): "ct_front_bn", id(net.cls_bn): "cls_bn", id(net.m_cls_bn): "m_cls_bn", id(net.keep_prob): "keep_prob"}

    class _Feed(object):
        def __init__(self, q):
            self.q = q

    class Session(object):
        def __init__(self, config=None):
            self.graph = "graph"

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def _one(self, f):
            if isinstance(f, _Feed):
                return np.zeros((BATCH, 256, 256, 4), np.float32) if f.q.endswith("data") else "fid"
            return f

        def run(self, fetches, feed_dict=None):
            feeds = {}
            for k, v in (feed_dict or {}).items():
                feeds[names[id(k)]] = "batch" if isinstance(v, np.ndarray) else v
            if isinstance(fetches, tuple) and len(fetches) == 2 and fetches[0] == "assign_lr":
                log.append({"op": "assign_lr", "value": fetches[1]})
                return None
            first = fetches[0] if isinstance(fetches, (list, tuple)) and fetches else fetches
            if first == "dis_optimizer" or first == "gen_optimizer":
                log.append({"op": first, "feeds": feeds})
                return None, me.learning_rate_node
            if isinstance(first, tuple) and first and first[0] == "clip_assign":
                log.append({"op": "clip_op", "n": len(fetches)})
                return [None] * len(fetches)
            if isinstance(first, tuple) and first and first[0] == "assign_lr":
                log.append({"op": "assign_lr", "value": first[1]})
                return None
            if isinstance(fetches, (list, tuple)):
                return [self._one(f) for f in fetches]
            return self._one(fetches)
    tf.Session = Session
    tf.ConfigProto = lambda: types.SimpleNamespace(gpu_options=types.SimpleNamespace(allow_growth=False))
    tf.summary = types.SimpleNamespace(FileWriter=lambda *a, **k: None)
    tf.train.Coordinator = lambda: types.SimpleNamespace(request_stop=lambda: None, join=lambda threads: None)
    tf.train.get_checkpoint_state = lambda path: None
    tf.train.start_queue_runners = lambda sess=None, coord=None, start=True: []

    class RMSPropOptimizer(object):
        n = 0

        def __init__(self, learning_rate=None, **kw):
            pass

        def minimize(self, loss, global_step=None, var_list=None):
            RMSPropOptimizer.n += 1
            return "dis_optimizer" if RMSPropOptimizer.n == 1 else "gen_optimizer"      # order of adversarial.py:643-651
    tf.train.RMSPropOptimizer = RMSPropOptimizer
    saved = (tf.Variable, tf.clip_by_value, getattr(tf, "assign", None))
    tf.Variable = lambda initial, trainable=True, name=None: float(initial) if not isinstance(initial, _Init) else saved[0](initial, trainable, name)
    tf.clip_by_value = lambda v, lo, hi, name=None: ("clip", v, float(lo), float(hi))
    tf.assign = lambda var, val: ("clip_assign", var) if hasattr(var, "rec") else ("assign_lr", float(val))

    class _Loss(object):
        def __add__(self, o):
            return self
        __radd__ = __add__
        __mul__ = __rmul__ = __add__
    net.dis_loss = net.dis_reg = net.ct_gen_loss = net.gen_reg = _Loss()

    class Me(adv.Trainer):
        def __init__(self):                                   # (the real one only stores arguments and opens the input queues)
            self.net = net
            self.num_cls = 5
            self.batch_size = BATCH
            self.opt_kwargs = {"learning_rate": 3e-4}
            self.train_config = {"restore_from_baseline": False, "copy_main": False, "clear_rms": False, "lr_update": True,
                                 "dis_interval": 1, "gen_interval": 1, "dis_sub_iter": 2, "gen_sub_iter": 1, "tag": "t",
                                 "iter_upd_interval": 2, "dis_sub_iter_inc": 1, "gen_sub_iter_inc": 0, "lr_decay_factor": 0.98,
                                 "checkpoint_space": 100, "training_iters": 5, "epochs": 1}
            self.lr_update_flag = self.train_config["lr_update"]
            self.ct_train_queue, self.ct_val_queue, self.mr_train_queue, self.mr_val_queue = "ct_train", "ct_val", "mr_train", "mr_val"

        def _initialize(self, training_iters, output_path):    # adversarial.py:658-705 minus the tensorboard summaries
            self.global_step = types.SimpleNamespace(eval=lambda: 0)
            self.dis_optimizer, self.gen_optimizer = adv.Trainer._get_optimizer(self, training_iters, self.global_step)
            return "init_glb", "init_loc"

        def next_batch(self, input_queue, **k):
            return _Feed(input_queue + ":data"), _Feed(input_queue + ":fid")

        def output_minibatch_stats(self, sess, writer, step, *a, **k):
            log.append({"op": "stats", "step": step, "detail": bool(k.get("detail", False))})
    me = Me()
    cfg_in = dict(me.train_config)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pnp_sched_")
    os.chdir(tmp)
    try:
        me.train(output_path="./out", restore=True, restored_path="./out", training_iters=5, epochs=1, dropout=0.75, display_step=10 ** 6)
    finally:
        os.chdir(cwd)
        tf.Variable, tf.clip_by_value = saved[0], saved[1]
        if saved[2] is not None:
            tf.assign = saved[2]
    return {"train_config": cfg_in, "train_args": {"training_iters": 5, "epochs": 1, "dropout": 0.75}, "events": log}


def pipeline_example():
    """the synthetic single-slice example both sides decode (tests rebuild it from this formula)"""
    i, j, c = np.meshgrid(np.arange(256), np.arange(256), np.arange(3), indexing="ij")
    data = (1000.0 * c + (i * 256 + j) % 997).astype(np.float32)
    label = ((i * 3 + j * 5 + c * 2) % 5).astype(np.float32)
    return data, label


def trace_input_pipeline(tf, adv):
    """Trainer.next_batch (adversarial.py:607-631) evaluated numerically on one parsed example: which bytes become the image
    and which channel of the stored label volume becomes the label map"""
    data, label = pipeline_example()
    feats = {"dsize_dim0": 256, "dsize_dim1": 256, "dsize_dim2": 3, "lsize_dim0": 256, "lsize_dim1": 256, "lsize_dim2": 3,
             "data_vol": data.tobytes(), "label_vol": label.tobytes()}
    seen = {}

    def parse_single_example(serialized, features):
        seen["feature_keys"] = sorted(features)
        return feats
    saved = {k: getattr(tf, k, None) for k in ("reshape", "slice", "concat", "cast")}
    tf.TFRecordReader = lambda: types.SimpleNamespace(read=lambda q: ("fid", "serialized"))
    tf.parse_single_example = parse_single_example
    tf.cast = lambda x, dtype: x
    tf.decode_raw = lambda b, dtype: np.frombuffer(b, dtype="<f4" if dtype == tf.float32 else None)
    tf.reshape = lambda x, shape: np.reshape(x, shape)
    tf.slice = lambda x, begin, size: x[tuple(slice(b, b + s_) for b, s_ in zip(begin, size))]
    tf.concat = lambda vals, axis, name=None: np.concatenate(vals, axis=axis)

    def shuffle_batch(tensors, batch_size, capacity, num_threads, min_after_dequeue):
        seen["shuffle_batch"] = {"batch_size": batch_size, "capacity": capacity, "num_threads": num_threads, "min_after_dequeue": min_after_dequeue}
        return [np.stack([t] * batch_size) if isinstance(t, np.ndarray) else [t] * batch_size for t in tensors]
    tf.train.shuffle_batch = shuffle_batch
    me = types.SimpleNamespace(batch_size=BATCH)
    pair, fid = adv.Trainer.next_batch(me, "queue")
    for k, v in saved.items():
        if v is not None:
            setattr(tf, k, v)
    assert pair.shape == (BATCH, 256, 256, 4)
    return {"feature_keys": seen["feature_keys"], "shuffle_batch": seen["shuffle_batch"], "pair_shape": list(pair.shape),
            "sample_rows": [0, 37, 128, 255], "sample_cols": [0, 41, 200, 255],
            "samples": pair[0][np.ix_([0, 37, 128, 255], [0, 41, 200, 255])].tolist(),
            "channel_sums": [float(pair[0, :, :, c].astype(np.float64).sum()) for c in range(4)]}


def trace_transplant(tf, adv, net, adv_names, baseline_names, adv_trainable):
    """adversarial.py:503-531 (restore, no_gan), :706-741 (_adapt_copy_weights, both modes), :743-765 (_load_batch_norm_weights)
    executed on the variable-name lists the reference ships under lists/ and on the traced variable tables"""
    def read(fn):
        with open(os.path.join(REF, "lists", fn)) as f:
            return [ln.split("\n")[0] for ln in f.readlines() if len(ln) >= 3]        # lib._read_lists, lib.py:7-20
    strip = lambda n: n.split(":")[0]
    pairs = []
    graph = types.SimpleNamespace(get_tensor_by_name=lambda n: ("tensor", n))
    tf.get_default_graph = lambda: graph
    saved_assign = getattr(tf, "assign", None)
    tf.assign = lambda dst, src: types.SimpleNamespace(eval=lambda: pairs.append((dst, src)))
    out = {}
    me = types.SimpleNamespace(mr_var_list=read("half_zip_mri_vars"), adapt_var_list=read("half_zip_ct_vars"),
                               old_bn_list=read("old_bn_list"), new_bn_list=read("pred_bn_list"))
    adv.Trainer._adapt_copy_weights(me, internal=False)
    out["adapt_copy"] = [[strip(d[1]), strip(s_[1])] for d, s_ in pairs]
    # internal = True: correspondence by creation order of the 'group*' and 'adapt*' variables
    del pairs[:]
    tf.contrib.framework.get_variables = lambda: [_VarObj({"name": n}) for n in adv_names]
    me2 = types.SimpleNamespace()
    try:
        adv.Trainer._adapt_copy_weights(me2, internal=True)
        out["adapt_copy_internal"] = [[strip(d.name), strip(s_.name)] for d, s_ in pairs]
    except ValueError as e:
        out["adapt_copy_internal_error"] = str(e)          # 'group' matches group_1..10: 153 names vs 101 adapt names
    del pairs[:]
    tf.train.get_checkpoint_state = lambda path: None
    tf.contrib.framework.load_variable = lambda path, name: ("ckpt", name)
    adv.Trainer._load_batch_norm_weights(me, "./ckpt")
    out["bn_copy"] = [[s_[1], strip(d[1])] for d, s_ in pairs]          # [old name in the baseline checkpoint, new variable]
    out["bn_copy_dict_len"] = len(me.copy_bn_dict)
    # restore(no_gan=True) from a baseline-segmenter checkpoint (its variables + Adam slots)
    ckpt = {n: None for n in baseline_names}
    ckpt.update({n + "/Adam": None for n in baseline_names if n.endswith("Variable") or "Variable_" in n})
    ckpt["beta1_power"] = None
    restored = []

    class Saver(object):
        def __init__(self, var_list=None):
            self.var_list = var_list

        def restore(self, sess, path):
            restored.append([strip(v.name) for v in self.var_list])
    tf.train.Saver = Saver
    tf.get_collection_ref = lambda name: []
    tf.global_variables = lambda: [_VarObj({"name": n}) for n in adv_names]
    tf.pywrap_tensorflow = types.SimpleNamespace(NewCheckpointReader=lambda path: types.SimpleNamespace(get_variable_to_shape_map=lambda: ckpt))
    rc = adv.Full_DRN.restore(net, "sess", "./ckpt/model", no_gan=True)
    assert rc == 0 and len(restored) == 1
    out["restore_no_gan"] = {"checkpoint_names": sorted(ckpt), "restored": restored[0]}

    # ---- the other branches of restore (adversarial.py:533-574) on GAN checkpoints.  The live graph also holds the RMSProp slots
    trainable = [n for n in adv_names if adv_trainable[n] and ("cls" in n or "adapt" in n)]
    slots = [n + sfx for n in trainable for sfx in ("/RMSProp", "/RMSProp_1")]
    graph_vars = list(adv_names) + slots
    tf.global_variables = lambda: [_VarObj({"name": n}) for n in graph_vars]
    tf.contrib.framework.get_variables = lambda: [_VarObj({"name": n}) for n in graph_vars]

    class StrictSaver(object):
        """tf.train.Saver.restore fails when a variable of its list is missing from the checkpoint"""

        def __init__(self, var_list=None):
            self.var_list = var_list

        def restore(self, sess, path):
            names = [strip(v.name) for v in self.var_list]
            missing = [n for n in names if n not in ckpt_now]
            if missing:
                raise KeyError("not found in checkpoint: %s" % missing[0])
            restored.append(names)
    tf.train.Saver = StrictSaver
    cases = {}
    full = {n: None for n in graph_vars}
    partial = {n: None for n in graph_vars if "mask_cls" not in n}          # e.g. a checkpoint written before the mask critic existed
    for label, ck, kw in (("full_default", full, {}), ("full_clear_rms", full, {"clear_rms": True}), ("partial_default", partial, {})):
        ckpt_now = ck
        tf.pywrap_tensorflow = types.SimpleNamespace(NewCheckpointReader=lambda path, ck=ck: types.SimpleNamespace(get_variable_to_shape_map=lambda: ck))
        del restored[:]
        adv.Full_DRN.restore(net, "sess", "./ckpt/model", **kw)
        cases[label] = {"checkpoint_names": sorted(ck), "restored": sorted(restored[-1]), "skip_kwd": list(net.network_config["restore_skip_kwd"])}
    out["restore_gan"] = cases
    if saved_assign is not None:
        tf.assign = saved_assign
    return out


def trace_segmenter_schedule(tf, mod, net):
    """source_segmenter.Trainer (__init__, _get_optimizer, _initialize, train, output_minibatch_stats, val_stats:
    source_segmenter.py:312-570) executed verbatim against a recording Session; only next_batch (TFRecord queue) is replaced."""
    log = []
    names = {id(net.x): "x", id(net.y): "y", id(net.main_bn): "main_bn", id(net.adapt_bn): "adapt_bn", id(net.keep_prob): "keep_prob"}

    class _Feed(object):
        def __init__(self, kind):
            self.kind = kind

        def value(self):
            return np.zeros((BATCH, 256, 256, 4), np.float32) if self.kind == "data" else [b"file:0"] * BATCH

        def eval(self):
            return self.value()

    class Session(object):
        graph = "graph"

        def __init__(self, config=None):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def run(self, fetches, feed_dict=None):
            feeds = {names[id(k)]: ("batch" if isinstance(v, np.ndarray) else v) for k, v in (feed_dict or {}).items()}
            fl = list(fetches) if isinstance(fetches, (list, tuple)) else [fetches]
            if fl and fl[0] == "adam_op":
                log.append({"op": "optimizer", "feeds": feeds})
                return None, 0.0, me.learning_rate_node
            if "scalar_summary" in fl and "train_images" in fl:
                log.append({"op": "minibatch_stats", "feeds": feeds})
            elif "scalar_summary" in fl and "val_images" in fl:
                log.append({"op": "val_stats", "feeds": feeds, "detail": len(fl) == 5})
            elif fl and isinstance(fl[0], tuple) and fl[0] and fl[0][0] == "assign_lr":
                log.append({"op": "assign_lr", "value": fl[0][1]})
                return None
            out = []
            for f in fl:
                if isinstance(f, _Feed):
                    out.append(f.value())
                elif isinstance(f, Sym) and f.src == "cm":
                    out.append(np.zeros((5, 5)))
                else:
                    out.append(0.0)
            return out if isinstance(fetches, (list, tuple)) else out[0]
    tf.Session = Session
    tf.ConfigProto = lambda: types.SimpleNamespace(gpu_options=types.SimpleNamespace(allow_growth=False))
    merged = {"n": 0}

    def merge(lst):
        merged["n"] += 1
        return {1: "scalar_summary", 2: "train_images", 3: "val_images"}[merged["n"]]
    writer = lambda *a, **k: types.SimpleNamespace(add_summary=lambda s_, step: None, flush=lambda: None)
    tf.summary = types.SimpleNamespace(scalar=lambda n, v: ("s", n), image=lambda n, v: ("i", n), merge=merge, FileWriter=writer)
    opt_log = []

    class AdamOptimizer(object):
        def __init__(self, learning_rate=None, **kw):
            self.cfg = {"kind": "AdamOptimizer", "learning_rate": learning_rate, "kwargs": dict(kw)}

        def minimize(self, loss, global_step=None, var_list=None):
            rec = dict(self.cfg)
            rec["objective_src"] = loss.src
            rec["var_list"] = None if var_list is None else len(var_list)
            opt_log.append(rec)
            return "adam_op"
    train = types.ModuleType("tensorflow.train")
    train.AdamOptimizer = AdamOptimizer
    train.string_input_producer = lambda lst, num_epochs=None, shuffle=True: "queue"
    train.Coordinator = lambda: types.SimpleNamespace(request_stop=lambda: None, join=lambda threads: None)
    train.get_checkpoint_state = lambda path: None
    train.start_queue_runners = lambda sess=None, coord=None, start=True: []
    tf.train = train
    saved_var = tf.Variable
    class _Scalar(float):
        def eval(self):
            return float(self)
    tf.Variable = lambda initial, trainable=True, name=None: (saved_var(initial, trainable, name) if isinstance(initial, _Init)
                                                               else (_Scalar(initial) if np.isscalar(initial) else "var"))
    tf.assign = lambda var, val: ("assign_lr", float(val))
    tf.global_variables_initializer = lambda: "init_glb"
    tf.variables_initializer = lambda v: "init_loc"
    tf.local_variables = lambda: []
    tf.trainable_variables = lambda: [_VarObj(G.vars[n]) for n in G.order if G.vars[n]["trainable"]]
    net.cost.src, net.regularizer_loss.src = "cost", "regularizer_loss"

    class Me(mod.Trainer):
        def next_batch(self, input_queue, **k):                 # source_segmenter.py:331-355: TFRecord queue plumbing
            return _Feed("data"), _Feed("fid")
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pnp_segsched_")
    os.chdir(tmp)
    try:
        # the arguments train_segmenter.py:60-75 passes, with a short loop
        me = Me(net, train_list=["a"], val_list=["b"], num_cls=5, batch_size=BATCH, opt_kwargs={"learning_rate": 1e-3},
                checkpoint_space=1500, optimizer="adam", lr_update_flag=False)
        me.train(output_path="./out", training_iters=7, epochs=1, restore=True, restored_path="./out")
    finally:
        os.chdir(cwd)
        tf.Variable = saved_var
    return {"optimizer": opt_log, "events": log, "train_args": {"training_iters": 7, "epochs": 1, "display_step_default": 5, "dropout_default": 0.75}}


def trace_source_segmenter(tf):
    """source_segmenter.py does not parse (SyntaxError at :611, inside Trainer.test_eval).  Everything before `class Trainer`
    (line 303) -- the module preamble and class Full_DRN in full -- is compiled and executed verbatim."""
    global REC
    path = os.path.join(REF, "source_segmenter.py")
    with open(path) as f:
        src = f.read()
    try:
        compile(src, path, "exec")
        syntax_error = None
    except SyntaxError as e:
        syntax_error = {"line": e.lineno, "msg": e.msg}
    lines = src.split("\n")
    cut = next(i for i, ln in enumerate(lines) if ln.startswith("class Trainer"))
    cut2 = next(i for i, ln in enumerate(lines) if ln.strip().startswith("def test_eval"))     # the method holding the syntax error
    head = "\n".join(lines[:cut2]) + "\n"            # module preamble, class Full_DRN, class Trainer up to (not incl.) test_eval
    REC = Recorder()
    Sym._n = 0
    G.reset()
    mod = types.ModuleType("pnp_reference_source_segmenter")
    mod.__file__ = path
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pnp_trace_")
    os.chdir(tmp)
    try:
        exec(compile(head, path, "exec"), mod.__dict__)
        REC.set_section("create_network")
        cost_kwargs = {"cross_flag": True, "miu_cross": 1.0, "dice_flag": True, "miu_dice": 1.0, "regularizer": 1e-4}
        # distinct trainable flags show which groups each flag governs (source_segmenter.py:91-209)
        net = mod.Full_DRN(channels=3, n_class=5, batch_size=BATCH, main_trainable=False, adapt_trainable=True, cost_kwargs=cost_kwargs)
        REC._close()
    finally:
        os.chdir(cwd)
    for ev in REC.events:
        ev.pop("_x", None)
        ev.pop("skip_from_conv_input_of", None)
    # source_segmenter.py:241-273 evaluated numerically on a small map (logits scaled so the 0.005 clip is active)
    rng = np.random.RandomState(11)
    logits = 4.0 * rng.standard_normal((2, 6, 5, 5))
    lab = rng.randint(0, 5, size=(2, 6, 5))
    lab[0, :3] = 0
    y = np.eye(5)[lab]
    me = types.SimpleNamespace(y=y, n_class=5)
    with _Num(tf):
        wce = float(mod.Full_DRN._softmax_weighted_loss(me, logits))
        dice = float(mod.Full_DRN._dice_loss_fun(me, logits))
    losses = {"logits": logits.tolist(), "labels": lab.tolist(), "weighted_loss": wce, "dice_loss": dice}
    schedule = trace_segmenter_schedule(tf, mod, net)
    return {"syntax_error": syntax_error, "executed_lines": cut, "executed_lines_trainer": cut2, "events": REC.events,
            "variables": [G.vars[n] for n in G.order], "losses_numeric": losses, "schedule": schedule,
            "conv_weights": [s.var["name"] for s in net.conv_weights],
            "ctor_args": {"main_trainable": False, "adapt_trainable": True}}


if __name__ == "__main__":
    main()
