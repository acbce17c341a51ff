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
