"""Generate tests/golden/reference_lib_vectors.npz by EXECUTING the reference's own Python files in this container.

What can run without TensorFlow: `ops.py` (PS / _phase_shift, ops.py:3-27) only uses six shape ops, and the metric helpers of
`lib.py` (`_label_decomp` :75-92, `_jaccard` :121-135, `_dice` :138-152) are plain numpy.  The two files are imported unmodified
from /root/reference with `tensorflow` replaced by a numpy shim that implements exactly those six ops with their documented
TF-1.4 semantics (reshape, transpose, split(value, num, axis), concat(values, axis), squeeze(x) = drop ALL size-1 dims,
expand_dims) and `nibabel` by an empty module (lib.py imports it at the top; nothing here calls it).

The outputs are therefore produced by the reference's code, not by a restatement; tests/test_reference_golden.py pins the
oracle (and the product's host helpers) to them, and the GPU parity tests pin the CUDA kernels to the oracle.

    python tests/golden/make_reference_lib_vectors.py          # needs /root/reference; the tests only read the .npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_lib_vectors.npz")


class _Shape(object):
    def __init__(self, shp):
        self._s = list(shp)

    def as_list(self):
        return list(self._s)


class T(np.ndarray):
    """numpy array that answers the one tensor method ops.py calls"""

    def get_shape(self):
        return _Shape(self.shape)


def _t(a):
    return np.asarray(a).view(T)


def _make_tf_shim():
    tf = types.ModuleType("tensorflow")
    tf.reshape = lambda x, shape: _t(np.reshape(np.asarray(x), tuple(shape)))
    tf.transpose = lambda x, perm: _t(np.transpose(np.asarray(x), tuple(perm)))
    tf.split = lambda value, num, axis: [_t(p) for p in np.split(np.asarray(value), num, axis=axis)]
    tf.concat = lambda values, axis: _t(np.concatenate([np.asarray(v) for v in values], axis=axis))
    tf.squeeze = lambda x: _t(np.squeeze(np.asarray(x)))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(np.asarray(x), axis))
    # arithmetic helpers for layers.pixel_wise_softmax_2 (layers.py:134-138) and lib._dice_eval (lib.py:96-110)
    tf.exp = lambda x: np.exp(np.asarray(x))
    tf.reduce_sum = lambda x, axis=None, keep_dims=False: np.sum(np.asarray(x), axis=axis, keepdims=keep_dims)
    tf.shape = lambda x: np.asarray(x).shape
    tf.stack = lambda vals: [int(v) for v in vals]
    tf.tile = lambda x, multiples: np.tile(np.asarray(x), [int(m) for m in multiples])
    tf.div = lambda a, b, name=None: np.asarray(a) / np.asarray(b)
    tf.clip_by_value = lambda x, lo, hi, name=None: np.clip(x, lo, hi)
    tf.one_hot = lambda idx, depth, axis=-1: np.eye(depth)[np.asarray(idx)]
    return tf


def _load(name):
    spec = importlib.util.spec_from_file_location("pnp_reference_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.modules["tensorflow"] = _make_tf_shim()
    sys.modules["nibabel"] = types.ModuleType("nibabel")
    ops = _load("ops")
    lib = _load("lib")
    layers = _load("layers")
    rng = np.random.RandomState(20260924)
    out = {}

    # ---- PS: (batch, a, b, r, n_channel).  The graphs use r = 8, n_channel = 5 (on 32x32 maps; 16x16 here to keep the file small); B == 1 takes the other branch.
    ps_cases = [(2, 4, 4, 2, 3), (3, 4, 6, 2, 2), (2, 16, 16, 8, 5), (1, 4, 4, 2, 3), (1, 8, 8, 4, 2), (4, 3, 5, 3, 1)]
    out["ps_cases"] = np.array(ps_cases, dtype=np.int64)
    # PS is a pure permutation: the input is arange(numel) (exact in float32), only the permuted index map is stored
    for i, (B, a, b, r, nc) in enumerate(ps_cases):
        n = B * a * b * r * r * nc
        assert n < (1 << 24)
        x = np.arange(n, dtype=np.float32).reshape(B, a, b, r * r * nc)
        y = np.asarray(ops.PS(_t(x), r, n_channel=nc, batch_size=B))
        assert y.shape == (B, a * r, b * r, nc), y.shape
        out["ps_perm_%d" % i] = y.astype(np.uint32)

    # ---- lib._label_decomp on [B, H, W] integer maps
    lab = rng.randint(0, 5, size=(3, 9, 7)).astype(np.int64)
    lab[0, :, :] = 0                                   # an all-background slice
    out["ld_labels"] = lab
    out["ld_onehot"] = lib._label_decomp(5, lab)

    # ---- lib._dice / lib._jaccard on confusion matrices, including empty classes (zero row AND column -> 0, not NaN)
    cms = []
    for k in range(4):
        cm = rng.randint(0, 50, size=(5, 5)).astype(np.float64)
        if k == 1:
            cm[3, :] = 0
            cm[:, 3] = 0
        if k == 2:
            cm[:] = 0
            cm[0, 0] = 17
        cms.append(cm)
    cms = np.stack(cms)
    out["cm"] = cms
    out["cm_dice"] = np.stack([lib._dice(c) for c in cms])
    out["cm_jaccard"] = np.stack([lib._jaccard(c) for c in cms])

    # ---- layers.pixel_wise_softmax_2: exp / sum WITHOUT max subtraction, clipped to +-1e15 (float64 here)
    sm_x = 3.0 * rng.standard_normal((2, 5, 4, 5))
    sm_x[0, 0, 0] = [700.0, 0.0, -700.0, 1.0, 2.0]          # large but finite in float64
    out["sm_x"] = sm_x
    out["sm_y"] = np.asarray(layers.pixel_wise_softmax_2(sm_x), dtype=np.float64)

    # ---- lib._dice_eval on an argmax prediction and one-hot labels
    de_pred = rng.randint(0, 5, size=(2, 6, 5))
    de_lab = rng.randint(0, 5, size=(2, 6, 5))
    de_lab[1] = 2                                             # classes absent from the labels of a slice
    mean, arr = lib._dice_eval(de_pred, np.eye(5)[de_lab], 5)
    out["de_pred"], out["de_lab"] = de_pred, de_lab
    out["de_mean"], out["de_arr"] = np.float64(mean), np.asarray(arr, dtype=np.float64)

    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
