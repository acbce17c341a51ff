"""Pins to outputs of the REFERENCE'S OWN CODE (tests/golden/reference_lib_vectors.npz, produced in the build container by
tests/golden/make_reference_lib_vectors.py, which executes /root/reference/ops.py and lib.py under a six-op numpy shim of
tensorflow).  CPU only: the oracle and the product's host helpers are compared with the golden vectors here; the GPU parity
tests compare the CUDA kernels with the same oracle functions (test_ops_gpu.py: PS / one-hot / confusion), which closes the chain
reference code -> oracle -> kernels for the phase shift (ops.py:3-27), the one-hot feed (lib.py:75-92) and the confusion-matrix
metrics (lib.py:121-152).  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "reference_lib_vectors.npz"))
PS_CASES = [tuple(int(v) for v in row) for row in GOLD["ps_cases"]]


def _arange_input(B, a, b, r, nc):
    n = B * a * b * r * r * nc
    return np.arange(n, dtype=np.float32).reshape(B, a, b, r * r * nc)


@pytest.mark.parametrize("i", range(len(PS_CASES)), ids=lambda i: "B%d_%dx%d_r%d_g%d" % PS_CASES[i])
def test_oracle_phase_shift_equals_the_reference_code(i):
    """ops.py:23-27 executed by the reference itself vs the oracle's literal emulation, closed-form index law and torch form
    (bit exact: PS is a permutation).  Covers the B >= 2 transposed sub-pixel order and the B == 1 branch."""
    from oracle import tf14_numpy as N, tf14_torch as T
    B, a, b, r, nc = PS_CASES[i]
    want = GOLD["ps_perm_%d" % i].astype(np.float32)
    x = _arange_input(B, a, b, r, nc)
    assert want.shape == (B, a * r, b * r, nc)
    assert sorted(want.reshape(-1).tolist()) == list(range(x.size))          # a permutation of the input
    np.testing.assert_array_equal(N.PS_literal(x, r, nc, B), want)
    if B >= 2 or a == b:       # the closed forms are defined where the product is (square maps for B == 1)
        np.testing.assert_array_equal(N.PS_closed_form(x, r, nc, B), want)
        np.testing.assert_array_equal(T.PS(torch.from_numpy(x), r, nc, B).numpy(), want)


def test_oracle_and_host_label_decomp_equal_the_reference_code():
    """lib.py:75-92"""
    from oracle import tf14_numpy as N
    from pnp_b200 import lib as P
    lab, want = GOLD["ld_labels"], GOLD["ld_onehot"]
    assert want.dtype == np.float32 and want.shape == lab.shape + (5,)
    np.testing.assert_array_equal(N.label_decomp(5, lab), want)
    got = P._label_decomp(5, lab)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, want)


def test_host_dice_and_jaccard_equal_the_reference_code():
    """lib.py:121-152, including classes that never occur (0, not NaN)"""
    from pnp_b200 import lib as P
    for cm, d, j in zip(GOLD["cm"], GOLD["cm_dice"], GOLD["cm_jaccard"]):
        np.testing.assert_allclose(P._dice(cm), d, rtol=0, atol=1e-15)
        np.testing.assert_allclose(P._jaccard(cm), j, rtol=0, atol=1e-15)
    assert np.isfinite(GOLD["cm_dice"]).all() and (GOLD["cm_dice"][1][3] == 0) and (GOLD["cm_jaccard"][1][3] == 0)


def test_oracle_dice_eval_consistent_with_reference_confusion_dice():
    """lib._dice_eval (lib.py:96-110, needs tf.one_hot -> restated in the oracle) must agree with the reference's
    confusion-matrix Dice (lib.py:138-152, executed) on the same prediction/label pair -- two reference formulas, one value."""
    from oracle import tf14_numpy as N
    rng = np.random.RandomState(3)
    pred = rng.randint(0, 5, size=(2, 8, 8))
    lab = rng.randint(0, 5, size=(2, 8, 8))
    cm = np.zeros((5, 5))
    for t, p in zip(lab.reshape(-1), pred.reshape(-1)):
        cm[t, p] += 1
    from pnp_b200 import lib as P        # pinned to the reference's _dice above
    mean, arr = N.dice_eval(pred, N.label_decomp(5, lab).astype(np.float64), 5)
    np.testing.assert_allclose(np.array(arr), P._dice(cm), rtol=1e-7)      # (the 1e-7 epsilon of lib.py:100)
    np.testing.assert_allclose(mean, P._dice(cm).mean(), rtol=1e-7)


def test_oracle_pixel_softmax_equals_the_reference_code():
    """layers.py:134-138 executed numerically: no max subtraction, clip to +-1e15"""
    from oracle import tf14_numpy as N, tf14_torch as T
    x, want = GOLD["sm_x"], GOLD["sm_y"]
    assert np.isfinite(want).all() and want[0, 0, 0, 0] == pytest.approx(1.0) and want[0, 0, 0, 2] == 0.0
    np.testing.assert_allclose(N.pixel_wise_softmax_2(x), want, rtol=1e-14, atol=0)
    np.testing.assert_allclose(T.pixel_wise_softmax_2(torch.from_numpy(x)).numpy(), want, rtol=1e-13, atol=0)


def test_oracle_dice_eval_equals_the_reference_code():
    """lib.py:96-110 executed numerically (tf.one_hot of the prediction, per-class 2*inse/(union+1e-7), mean over classes)"""
    from oracle import tf14_numpy as N, tf14_torch as T
    pred, lab = GOLD["de_pred"], GOLD["de_lab"]
    y = np.eye(5)[lab]
    mean, arr = N.dice_eval(pred, y, 5)
    np.testing.assert_allclose(np.array(arr), GOLD["de_arr"], rtol=1e-14)
    assert mean == pytest.approx(float(GOLD["de_mean"]), rel=1e-14)
    mt, at = T.dice_eval(torch.from_numpy(pred), torch.from_numpy(y), 5)
    np.testing.assert_allclose(np.array([float(a) for a in at]), GOLD["de_arr"], rtol=1e-12)
    assert float(mt) == pytest.approx(float(GOLD["de_mean"]), rel=1e-12)


@pytest.mark.parametrize("B", [2, 1, 3])
def test_oracle_discriminator_input_equals_the_executed_reference_head(B):
    """adversarial.py:320-335 executed numerically (tests/golden/make_reference_disc_input_vectors.py: the head of `create_classifier`
    run unmodified up to its first `residual_block` call, with the reference's own PS and simple_concat2d): the 32-channel tensor the
    feature discriminator sees -- PS'ed taps, the x3 tile, logits, float(argmax) with the first-index tie rule -- bit for bit; B = 1
    takes the transposed sub-pixel branch of ops.py:11-20"""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_reference_disc_input_vectors as gen
    from oracle.pnp_graphs import disc_input
    gold = np.load(os.path.join(HERE, "golden", "reference_disc_input_vectors.npz"))
    c4, c6, b7, c9, logits = gen.make_inputs(B, int(gold["seed_B%d" % B]))
    want = gold["input_comp_B%d" % B]
    got = disc_input(*[torch.from_numpy(t) for t in (c4, c6, b7, c9, logits)], B).numpy()
    assert got.shape == want.shape == (B, 16, 16, 32) and got.dtype == np.float32
    assert np.array_equal(got, want)
    assert want[0, 0, 0, 31] == 1.0                                        # the planted tie [1, 3, 3, 0, -1] -> first maximum
    # channel map of KAT 8 on the executed tensor: 0-5 = (c4 g0, c4 g1) x 3, 26-30 logits, 31 argmax
    assert np.array_equal(want[..., 0:2], want[..., 2:4]) and np.array_equal(want[..., 0:2], want[..., 4:6])
    assert np.array_equal(want[..., 26:31], logits) and np.array_equal(want[..., 31], logits.argmax(-1).astype(np.float32))
