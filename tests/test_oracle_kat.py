"""Known-answer tests that pin the oracle (oracle/) -- hand-derived from the reference's call sites and the
published TF-1.4 op semantics (SURVEY 8c list), plus cross-checks between the two independent oracle forms
(naive numpy loops vs torch-CPU).  The reference ships no tests or golden vectors and TF-1.4 cannot run here:
parity is UNPINNED by the reference itself; these KATs are what stands in for it.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import tf14_numpy as N
from oracle import tf14_torch as T
from oracle import pnp_graphs as G


def test_kat01_phase_shift_index_law_vs_literal_emulation():
    """(1) ops.py:3-27 executed literally in numpy pins the closed-form law, for B>=2 and the B==1 branches"""
    rng = np.random.RandomState(0)
    for B in (1, 2, 3):
        for G_ in (1, 3):
            r = 4
            # the reference's batch_size==1 branches (reshape + final transpose) are only shape-consistent for
            # square maps -- which is all the reference ever feeds (32x32); B>=2 is checked on a ragged map
            a, b = (4, 4) if B == 1 else (3, 5)
            X = rng.randn(B, a, b, G_ * r * r).astype(np.float32)
            lit = N.PS_literal(X, r, G_, B)
            assert lit.shape == (B, a * r, b * r, G_)
            assert np.array_equal(lit, N.PS_closed_form(X, r, G_, B))
            assert np.array_equal(lit, T.PS(torch.from_numpy(X), r, G_, B).numpy())
    # explicit spot value, B>=2: out[n, i*r+q, j*r+p, g] = X[n,i,j,g*r*r+p*r+q]
    # (1-pixel maps hit tf.squeeze's all-axes behaviour, so the spot check uses a 2x2 map and looks at pixel (0,0))
    X = np.zeros((2, 2, 2, 4), dtype=np.float32)
    X[:, 0, 0, :] = np.arange(4, dtype=np.float32)
    out = N.PS_literal(X, 2, 1, 2)
    assert out[0, :2, :2, 0].tolist() == [[0.0, 2.0], [1.0, 3.0]]    # (q=row, p=col): X[p*2+q]
    out1 = N.PS_literal(X[:1], 2, 1, 1)
    assert out1[0, :2, :2, 0].tolist() == [[0.0, 1.0], [2.0, 3.0]]   # B==1: standard order


def test_kat02_same_padding_offsets():
    """(2) TF SAME: out=ceil(in/s), total=max((out-1)s+k_eff-in,0), before=total//2"""
    assert N.same_pad(256, 3, 2) == (0, 1)
    assert N.same_pad(128, 5, 2) == (1, 2)
    assert N.same_pad(16, 5, 4) == (0, 1)
    assert N.same_pad(128, 5, 4) == (0, 1)
    assert N.same_pad(32, 5, 4) == (0, 1)
    assert N.same_pad(32, 3, 1) == (1, 1)
    assert N.same_pad(32, 3, 1, 2) == (2, 2)      # atrous rate 2: effective 5x5
    assert N.same_pad(256, 5, 1) == (2, 2)


def test_kat03_symmetric_pad():
    """(3) tf.pad SYMMETRIC of [a b c] by 2 -> [b a a b c c b] (edge included)"""
    x = np.array([1.0, 2.0, 3.0]).reshape(1, 1, 3, 1).repeat(3, axis=1)
    p = N.symmetric_pad(x, 2)
    assert p[0, 2, :, 0].tolist() == [2, 1, 1, 2, 3, 3, 2]
    assert np.array_equal(p, np.pad(x, [(0, 0), (2, 2), (2, 2), (0, 0)], mode="symmetric"))
    t = T.symmetric_pad(torch.from_numpy(x), 2).numpy()
    assert np.array_equal(t, p)


def test_kat04_symmetric_strided_output_sizes():
    """(4) cls_6: 4x4 -(mirror 1)-> 6x6 -VALID s2-> 2x2 ; m_cls_4: 8x8 -> 12x12 -s4-> 2x2"""
    y = N.conv2d(np.ones((1, 4, 4, 2)), np.ones((3, 3, 2, 3)), stride=2, padding="SYMMETRIC")
    assert y.shape == (1, 2, 2, 3) and np.allclose(y, 18.0)
    y = N.conv2d(np.ones((1, 8, 8, 1)), np.ones((5, 5, 1, 1)), stride=4, padding="SYMMETRIC")
    assert y.shape == (1, 2, 2, 1) and np.allclose(y, 25.0)


def test_kat05_batch_norm_train_and_moving_update():
    """(5) biased var to normalise, unbiased into moving var, eps 1e-3, decay .9, default init"""
    x = np.array([1.0, 2.0, 3.0, 6.0]).reshape(4, 1, 1, 1)
    y, mm, mv = N.batch_norm(x, np.ones(1), np.zeros(1), np.zeros(1), np.ones(1), True)
    mean, var = 3.0, (4 + 1 + 0 + 9) / 4.0
    assert np.allclose(y.ravel(), (x.ravel() - mean) / math.sqrt(var + 1e-3))
    assert np.allclose(mm, 0.1 * mean)
    assert np.allclose(mv, 0.9 * 1.0 + 0.1 * var * 4 / 3)
    y2, mm2, mv2 = N.batch_norm(x, np.ones(1), np.zeros(1), mm, mv, False)
    assert np.allclose(y2.ravel(), (x.ravel() - mm) / np.sqrt(mv + 1e-3)) and mm2 is mm and mv2 is mv
    bn = T.BNState(1, torch.float64)
    yt = T.batch_norm(torch.from_numpy(x), bn, True)
    assert np.allclose(yt.detach().numpy(), y) and np.allclose(bn.moving_mean.numpy(), mm) and np.allclose(bn.moving_var.numpy(), mv)


def test_kat06_leaky_relu_alpha():
    """(6) tf.nn.leaky_relu default alpha = 0.2"""
    assert N.leaky_relu(np.array([-5.0, 3.0])).tolist() == [-1.0, 3.0]
    assert T.act(torch.tensor([-5.0, 3.0]), True).tolist() == [-1.0, 3.0]
    assert T.act(torch.tensor([-5.0, 3.0]), False).tolist() == [0.0, 3.0]


def test_kat07_inc_dim_skip_layout():
    """(7) channels [C/2 zeros | x | C/2 zeros]"""
    x = np.arange(1, 5, dtype=np.float64).reshape(1, 1, 1, 4)
    assert N.channel_pad_skip(x).ravel().tolist() == [0, 0, 1, 2, 3, 4, 0, 0]
    assert T.channel_pad_skip(torch.from_numpy(x)).numpy().ravel().tolist() == [0, 0, 1, 2, 3, 4, 0, 0]


def test_kat08_discriminator_input_channel_order():
    """(8) 0-5 = (c4 g0, g1) x3 ; 6-9 c6 ; 10-17 b7 ; 18-25 c9 ; 26-30 logits ; 31 argmax"""
    B = 2
    ws, bns = G.OracleAdversarial.layout()
    P = G.init_numpy_params(ws, bns, 0, 0.05)
    o = G.OracleAdversarial(P, B, critic_keep_prob=1.0)
    mk = lambda c, v: torch.full((B, 32, 32, c), float(v))
    logits = torch.zeros(B, 256, 256, 5)
    logits[..., 3] = 1.0
    with torch.no_grad():
        _, d_in = o.classifier(mk(128, 4), mk(256, 6), mk(512, 7), mk(512, 9), logits)
    assert d_in.shape == (B, 256, 256, 32)
    v = d_in[0, 17, 101]
    assert v[:6].tolist() == [4.0] * 6 and v[6:10].tolist() == [6.0] * 4
    assert v[10:18].tolist() == [7.0] * 8 and v[18:26].tolist() == [9.0] * 8
    assert v[26:31].tolist() == [0, 0, 0, 1.0, 0] and v[31].item() == 3.0


def test_kat09_weighted_ce_and_dice_toy():
    """(9) 2 pixels x 2 classes, by hand"""
    logits = np.array([[0.0, 0.0], [math.log(3.0), 0.0]]).reshape(1, 2, 1, 2)     # p = [.5,.5], [.75,.25]
    y = np.array([[1.0, 0.0], [0.0, 1.0]]).reshape(1, 2, 1, 2)
    # w0 = 1 - 1/2 = .5, w1 = .5 ; raw = -.5*log(.5) (pixel 0) , -.5*log(.25) (pixel 1) ; mean over 2 pixels
    wce = (-0.5 * math.log(0.5) - 0.5 * math.log(0.25)) / 2
    assert abs(N.softmax_weighted_loss(logits, y) - wce) < 1e-12
    # clip at 0.005: a probability of 1e-4 is treated as 0.005
    l2 = np.array([[0.0, 0.0], [0.0, math.log(9999.0)]]).reshape(1, 2, 1, 2)      # pixel 1: p0 = 1e-4
    y2 = np.array([[1.0, 0.0], [1.0, 0.0]]).reshape(1, 2, 1, 2)
    exp = (-0.0 * 1) + 0  # w0 = 1 - 2/2 = 0 -> everything weighted by zero
    assert abs(N.softmax_weighted_loss(l2, y2) - exp) < 1e-12
    inse0, l0, r0 = 0.5 * 1 + 0.75 * 0, 0.25 + 0.5625, 1.0
    inse1, l1, r1 = 0.5 * 0 + 0.25 * 1, 0.25 + 0.0625, 1.0
    dice = -(2 * inse0 / (l0 + r0 + 1e-7) + 2 * inse1 / (l1 + r1 + 1e-7)) / 2
    assert abs(N.dice_loss(logits, y) - dice) < 1e-12
    lt, yt = torch.from_numpy(logits), torch.from_numpy(y)
    assert abs(float(T.softmax_weighted_loss(lt, yt)) - wce) < 1e-12 and abs(float(T.dice_loss(lt, yt)) - dice) < 1e-12


def test_kat10_tf_adam_and_rmsprop_scalar_steps():
    """(10) TF Adam: lr_t = lr sqrt(1-b2^t)/(1-b1^t), eps added to sqrt(v) ('epsilon hat');
    TF RMSProp: ms0 = 1, eps inside the sqrt, momentum 0"""
    th, m, v = N.adam_step(1.0, 0.5, 0.0, 0.0, 1)
    m1, v1 = 0.1 * 0.5, 0.001 * 0.25
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(th - (1.0 - lr_t * m1 / (math.sqrt(v1) + 1e-8))) < 1e-15 and abs(m - m1) < 1e-15 and abs(v - v1) < 1e-15
    th, ms, mom = N.rmsprop_step(1.0, 0.5, 1.0, 0.0)
    ms1 = 0.9 * 1.0 + 0.1 * 0.25
    assert abs(ms - ms1) < 1e-15 and abs(th - (1.0 - 3e-4 * 0.5 / math.sqrt(ms1 + 1e-10))) < 1e-15
    p = torch.tensor([1.0], dtype=torch.float64)
    o = T.TFAdam([p])
    o.step([torch.tensor([0.5], dtype=torch.float64)])
    assert abs(float(p) - (1.0 - lr_t * m1 / (math.sqrt(v1) + 1e-8))) < 1e-15
    q = torch.tensor([1.0], dtype=torch.float64)
    o = T.TFRMSProp([q])
    o.step([torch.tensor([0.5], dtype=torch.float64)])
    assert abs(float(q) - (1.0 - 3e-4 * 0.5 / math.sqrt(ms1 + 1e-10))) < 1e-15


def test_kat11_l2_loss():
    """(11) tf.nn.l2_loss = sum(w^2)/2"""
    assert N.l2_loss([1.0, 2.0, 3.0]) == 7.0
    assert float(T.l2_loss(torch.tensor([1.0, 2.0, 3.0]))) == 7.0


def test_kat12_dice_eval_counts_background():
    """(12) hard Dice averaged over all 5 classes, background included"""
    lab = np.zeros((1, 2, 2), np.int64)
    lab[0, 0, 0] = 1
    pred = np.zeros((1, 2, 2), np.int64)
    y = N.label_decomp(5, lab)
    d, arr = N.dice_eval(pred, y.astype(np.float64), 5)
    # class 0: inse 3, union 4+3 -> 6/7 ; class 1: 0/(0+1) ; classes 2-4: 0/(0+eps) = 0
    assert abs(arr[0] - 6 / (7 + 1e-7)) < 1e-12 and arr[1] == 0 and abs(d - arr[0] / 5) < 1e-12
    dt, _ = T.dice_eval(torch.from_numpy(pred), torch.from_numpy(y).double(), 5)
    assert abs(float(dt) - d) < 1e-12


@pytest.mark.parametrize("case", [(2, 7, 9, 3, 4, 3, 1, 1, "SAME"), (1, 8, 8, 2, 3, 3, 2, 1, "SAME"), (1, 9, 9, 2, 2, 5, 2, 1, "SAME"),
                                  (1, 16, 16, 1, 2, 5, 4, 1, "SAME"), (2, 6, 6, 2, 3, 3, 1, 2, "SAME"), (1, 6, 6, 2, 2, 3, 1, 1, "SYMMETRIC"),
                                  (1, 4, 4, 2, 2, 3, 2, 1, "SYMMETRIC"), (1, 8, 8, 1, 2, 5, 4, 1, "SYMMETRIC")])
def test_conv_forms_agree(case):
    """naive-loop conv (with TF's asymmetric SAME / mirror pad) == torch-CPU form, fp64"""
    B, H, W, Cin, Cout, k, s, d, pad = case
    rng = np.random.RandomState(1)
    x = rng.randn(B, H, W, Cin)
    w = rng.randn(k, k, Cin, Cout)
    a = N.conv2d(x, w, s, d, pad)
    b = T.conv2d_raw(torch.from_numpy(x), torch.from_numpy(w), s, d, pad).numpy()
    assert a.shape == b.shape and np.allclose(a, b, atol=1e-10)


def test_truncated_normal_redraws():
    rng = np.random.RandomState(0)
    t = N.truncated_normal(rng, (10000,), 0.1)
    assert np.abs(t).max() <= 0.2 + 1e-7 and 0.08 < t.std() < 0.095


def test_oracle_layouts_match_reference_variable_lists():
    """the variable names the oracle (and the product) use == the reference's checkpoint naming contract.
    Golden name lists derived from lists/half_zip_*_vars and lists/*_bn_list are committed under tests/golden/."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_var_names.json")))
    ws, bns = G.OracleAdversarial.layout()
    names = set(n for n, _ in ws) | set(n + "/" + l for n, _ in bns for l in ("beta", "gamma", "moving_mean", "moving_variance"))
    for key in ("half_zip_mri_vars", "half_zip_ct_vars"):
        missing = [n for n in gold[key] if n not in names]
        assert not missing, (key, missing[:5])
    leafs = set(n.split("/", 1)[1] for n in names if "/" in n)
    assert not [n for n in gold["pred_bn_list"] if n not in leafs]
    ws, bns = G.OracleSegmenter.layout()
    seg_names = set(n + "/" + l for n, _ in bns for l in ("beta", "gamma", "moving_mean", "moving_variance"))
    assert sorted(seg_names) == sorted(gold["old_bn_list"])
    assert sum(int(np.prod(s)) for _, s in ws) == 39302456       # SURVEY Appendix A.2 total


def test_same_padding_agrees_with_an_independent_port_of_tensorflows_rule():
    """The SAME-padding offsets are defined by TensorFlow's kernels, which cannot run here.  Independent evidence: the
    `transformers` package in this image carries its own port of TF's rule for the MobileNet checkpoints
    (`apply_tf_padding`, citing tensorflow.org 'notes on padding').  The oracle's same_pad -- and the product's host copy --
    must place the same zeros for every (size, kernel, stride) the graphs use and a sweep around them (dilation 1)."""
    import torch
    tr = pytest.importorskip("transformers.models.mobilenet_v1.modeling_mobilenet_v1")
    from oracle import tf14_numpy as N
    from pnp_b200 import functional as F
    for n in list(range(1, 40)) + [64, 128, 255, 256, 257]:
        for k in (1, 3, 5, 7):
            for s in (1, 2, 3, 4):
                conv = torch.nn.Conv2d(1, 1, k, stride=s)
                x = torch.ones(1, 1, n, n)
                y = tr.apply_tf_padding(x, conv)
                total = y.shape[-1] - n
                before = int((y[0, 0].sum(0) > 0).float().argmax())                    # zero columns placed before the data
                assert N.same_pad(n, k, s) == (before, total - before), (n, k, s)
                assert tuple(F.same_pad(n, k, s)) == (before, total - before), (n, k, s)
                assert y.shape[-1] >= k and (y.shape[-1] - k) // s + 1 == -(-n // s)            # SAME output size ceil(n / s)


def test_kat13_pooling_same_geometry_for_any_n():
    """tf.nn.max_pool / avg_pool, ksize = strides = n, 'SAME' (layers.py:102-106): out = ceil(in / n), pad_before = (out * n - in) // 2,
    padding never wins a max and is not counted by the average.  Hand-derived on arange maps; both oracle forms."""
    import torch
    from oracle import tf14_torch as T
    x = np.arange(25, dtype=np.float64).reshape(1, 5, 5, 1)              # 5 -> 3 windows: rows [0,2) [2,4) [4,5)
    assert N.pool_same(x, 2)[0, :, :, 0].tolist() == [[6, 8, 9], [16, 18, 19], [21, 23, 24]]
    assert N.pool_same(x, 2, avg=True)[0, :, :, 0].tolist() == [[3, 5, 6.5], [13, 15, 16.5], [20.5, 22.5, 24]]
    x = np.arange(16, dtype=np.float64).reshape(1, 4, 4, 1)              # n = 3: pad_needed 2, one row of padding BEFORE: rows [0,2) [2,4)
    assert N.pool_same(x, 3)[0, :, :, 0].tolist() == [[5, 7], [13, 15]]
    assert N.pool_same(x, 3, avg=True)[0, :, :, 0].tolist() == [[2.5, 4.5], [10.5, 12.5]]
    assert np.array_equal(N.pool_same(x, 1), x) and np.array_equal(N.pool_same(x, 2), N.max_pool2x2(x))
    rng = np.random.RandomState(0)
    for (H, W, n) in [(5, 7, 2), (8, 12, 2), (9, 9, 4), (4, 4, 3), (6, 5, 1), (3, 10, 5)]:
        x = rng.standard_normal((2, H, W, 3))
        for avg in (False, True):
            a, b = N.pool_same(x, n, avg), T.pool_same(torch.from_numpy(x), n, avg).numpy()
            assert a.shape == b.shape == (2, -(-H // n), -(-W // n), 3) and np.abs(a - b).max() < 1e-12
