"""Generate tests/golden/reference_disc_input_vectors.npz by EXECUTING the head of the reference's `Full_DRN.create_classifier`
(adversarial.py:320-335) numerically.

The method's definition is pulled out of /root/reference/adversarial.py with `ast` and executed unmodified.  Its first ten statements
build the 32-channel discriminator input from four feature maps and the logits with `PS` (the reference's own ops.py, executed under
the numpy shim of make_reference_lib_vectors.py), `tf.tile`, `simple_concat2d` (the reference's own layers.py), `tf.argmax`, `tf.cast`,
`tf.expand_dims`; the first `residual_block(input_comp, ...)` call is where the network proper starts, so `residual_block` is replaced
by a function that hands its first argument back through an exception.  Nothing of the reference is copied into the repository.

Spatial size 2 x 2 feature maps (16 x 16 after PS; ops.PS's structure does not depend on the size), B = 2 (the batch >= 2 branch the
training graphs use) and B = 1 (the transposed special case of ops.py:11-20).  tests/test_reference_golden.py holds the oracle's
`disc_input` to the stored tensors bit-exactly; the GPU gather kernel is held to the oracle by tests/test_ops_gpu.py.

    python tests/golden/make_reference_disc_input_vectors.py       # needs /root/reference; the tests only read the .npz
"""
import ast
import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "reference_disc_input_vectors.npz")


def make_inputs(B, seed, a=2):
    """seeded feature maps / logits (shared with the test): c4 [B,a,a,128], c6 [B,a,a,256], b7, c9 [B,a,a,512], logits [B,8a,8a,5]"""
    rng = np.random.RandomState(seed)
    c4, c6, b7, c9 = [rng.standard_normal((B, a, a, c)).astype(np.float32) for c in (128, 256, 512, 512)]
    logits = rng.standard_normal((B, 8 * a, 8 * a, 5)).astype(np.float32)
    logits[0, 0, 0] = [1.0, 3.0, 3.0, 0.0, -1.0]             # a tie: tf.argmax returns the first maximal index
    return c4, c6, b7, c9, logits


class _Captured(Exception):
    def __init__(self, value):
        self.value = value


def main():
    sys.path.insert(0, HERE)
    import make_reference_lib_vectors as shim
    tf = shim._make_tf_shim()
    tf.variable_scope = lambda name: contextlib.nullcontext(name)
    tf.float32 = np.float32
    tf.argmax = lambda x, axis: np.argmax(np.asarray(x), axis=axis)
    tf.cast = lambda x, dtype: shim._t(np.asarray(x).astype(dtype))
    tf.equal = lambda a, b: np.array_equal(a, b)
    _tile = tf.tile
    tf.tile = lambda x, multiples: shim._t(_tile(x, multiples))
    sys.modules["tensorflow"] = tf
    sys.modules["nibabel"] = types.ModuleType("nibabel")

    def load(name):
        spec = importlib.util.spec_from_file_location("pnp_reference_" + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    ops, layers = load("ops"), load("layers")
    tree = ast.parse(open(os.path.join(REF, "adversarial.py")).read())
    (cls,) = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Full_DRN"]
    (fn,) = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "create_classifier"]

    def residual_block(x, *a, **k):
        raise _Captured(np.asarray(x))
    ns = {"tf": tf, "PS": ops.PS, "simple_concat2d": layers.simple_concat2d, "sharable_weight_variable": lambda **k: None,
          "residual_block": residual_block}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REF, "adversarial.py"), "exec"), ns)
    out = {}
    for B, seed in ((2, 11), (1, 12), (3, 13)):
        c4, c6, b7, c9, logits = make_inputs(B, seed)
        me = types.SimpleNamespace(batch_size=B)
        try:
            ns["create_classifier"](me, shim._t(c4), shim._t(c6), shim._t(b7), shim._t(c9), shim._t(logits))
            raise AssertionError("create_classifier returned without reaching residual_block")
        except _Captured as c:
            got = c.value
        assert got.shape == (B, 16, 16, 32), got.shape
        out["input_comp_B%d" % B] = got.astype(np.float32)
        out["seed_B%d" % B] = np.int64(seed)
        print("B=%d: input_comp %s, argmax channel at the tie = %g" % (B, got.shape, got[0, 0, 0, 31]))
    np.savez_compressed(OUT, **out)
    print("wrote %s (%.0f KB)" % (OUT, os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
