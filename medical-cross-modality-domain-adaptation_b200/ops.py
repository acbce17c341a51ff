"""Phase-shift (pixel-shuffle x r) upsampling with the reference's exact sub-pixel ordering
(ops.py:3-27), as one gather kernel instead of ~130 reshape/split/concat TF ops per call.

Index law (pinned against a literal emulation of the reference's op sequence in
tests/test_oracle_kat.py):
    batch_size >= 2 : out[n, i*r+q, j*r+p, g] = X[n, i, j, g*r*r + p*r + q]
    batch_size == 1 : out[n, i*r+p, j*r+q, g] = X[n, i, j, g*r*r + p*r + q]
(the reference's batch_size==1 branches yield the transposed order, ops.py:11-20).
"""
from . import functional as F


def _check(X, batch_size):
    if X.shape[0] != batch_size:
        raise ValueError("PS: tensor batch %d != batch_size %d (the reference reshapes with a static batch)"
                         % (X.shape[0], batch_size))
    if X.shape[1] < 2 or X.shape[2] < 2:
        raise ValueError("PS: 1-pixel maps hit tf.squeeze's all-axes behaviour in the reference (ops.py:10,17); unsupported")
    if batch_size == 1 and X.shape[1] != X.shape[2]:
        raise ValueError("PS: the reference's batch_size==1 branch (ops.py:11-20) is only shape-consistent for square maps")


def _phase_shift(I, r, batch_size=10):
    """ops.py:3-21: one group of r*r channels -> one channel upsampled r times"""
    _check(I, batch_size)
    return F.phase_shift(I, r, 1, batch_size)


def PS(X, r, n_channel=8, batch_size=10):
    """ops.py:23-27: split channels into n_channel groups, phase-shift each, concat"""
    _check(X, batch_size)
    return F.phase_shift(X, r, n_channel, batch_size)
