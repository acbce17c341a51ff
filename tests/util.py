"""shared helpers for the parity tests: CUDA product (pnp_b200, through the C-ABI) vs CPU oracle"""
import numpy as np
import torch


def rel_err(a, b):
    """max |a-b| / max|b| (the SURVEY 8d definition: error relative to the largest reference magnitude)"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def check(name, got, ref, tol):
    e = rel_err(got, ref)
    print("  %-40s rel_err %.3e (tol %.1e) shape %s" % (name, e, tol, tuple(ref.shape)))
    assert tuple(got.shape) == tuple(ref.shape), "%s: shape %s vs %s" % (name, tuple(got.shape), tuple(ref.shape))
    assert np.isfinite(e) and e <= tol, "%s: rel err %.3e > %.1e" % (name, e, tol)
    return e


def randn(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale).to(dtype)


def l2_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def check_grad(name, got, ref, tol_max, tol_l2):
    """gradient check in two norms.  Train-mode BN backward subtracts the per-channel mean of the incoming gradient, which
    amplifies upstream rounding (an fp32 CPU reference deviates from fp64 by up to ~1e-2 in max-norm on these graphs), so
    gradients are held to a max-norm AND a (much tighter) relative-L2 bound."""
    e, l = rel_err(got, ref), l2_err(got, ref)
    print("  %-40s rel_err %.3e (tol %.1e)  l2 %.3e (tol %.1e)" % (name, e, tol_max, l, tol_l2))
    assert tuple(got.shape) == tuple(ref.shape)
    assert np.isfinite(e) and e <= tol_max and l <= tol_l2, "%s: max %.3e l2 %.3e" % (name, e, l)


def compare_with_opencv_vectors(tag, logits, tol_rel=1e-3, min_agree=0.999):
    """`logits` [B,256,256,5] (torch / numpy) against tests/golden/opencv_reference_graph_vectors.npz: the same graph -- the reference's,
    from its recorded trace -- executed by OpenCV's TensorFlow importer (tests/golden/make_opencv_reference_vectors.py).
    Returns (max error at the stored positions relative to the largest |logit|, argmax agreement over the full maps)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "opencv_reference_graph_vectors.npz"))
    lg = logits.detach().float().cpu().numpy() if hasattr(logits, "detach") else np.asarray(logits, np.float32)
    B = int(gold["batch"])
    assert lg.shape == (B, 256, 256, 5), lg.shape
    pos = gold["positions"]
    flat = lg.reshape(B, 256 * 256, 5)
    got = np.stack([flat[b][pos[b]] for b in range(B)])
    err = float(np.abs(got - gold[tag + "_logits_at_positions"]).max() / float(gold[tag + "_max_abs_logit"]))
    agree = float((lg.argmax(-1) == gold[tag + "_argmax"]).mean())
    print("  %-12s vs OpenCV-executed reference graph: logits max rel err %.3e (tol %.1e), argmax agreement %.6f" % (tag, err, tol_rel, agree))
    assert np.isfinite(err) and err <= tol_rel and agree >= min_agree, (tag, err, agree)
    # north star: hard Dice (lib.py:96-110, background included) on held-out synthetic labels (seed 7777) within 1e-3 of the reference --
    # here the reference IS the third-party execution of the reference graph
    from oracle.pnp_graphs import synthetic_labels
    lab = np.asarray(synthetic_labels(B, 7777))

    def hard_dice(pred):
        return np.array([2.0 * np.sum((pred == c) & (lab == c)) / (np.sum(pred == c) + np.sum(lab == c) + 1e-7) for c in range(5)])
    d_got, d_ref = hard_dice(lg.argmax(-1)), hard_dice(gold[tag + "_argmax"].astype(np.int64))
    print("  %-12s hard Dice on held-out labels: ours %s, OpenCV %s" % (tag, np.round(d_got, 5), np.round(d_ref, 5)))
    assert np.abs(d_got - d_ref).max() <= 1e-3 and abs(d_got.mean() - d_ref.mean()) <= 1e-3
    return err, agree
