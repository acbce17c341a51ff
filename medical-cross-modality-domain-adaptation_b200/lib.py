"""Helpers of the reference's lib.py that sit on the hot path (one-hot feed, Dice evaluation) plus the
confusion-matrix Dice/Jaccard used for console statistics.  NIfTI I/O (lib.py:31-72) is out of scope
(`nibabel` is absent and the MMWHS volumes are not available offline)."""
import numpy as np
import torch

from . import functional as F


def _read_lists(fid):
    """lib.py:7-20: read a text file of paths, one per line (lines shorter than 3 chars skipped)."""
    with open(fid, 'r') as fd:
        lines = fd.readlines()
    return [ln.split('\n')[0] for ln in lines if len(ln) >= 3]


def _label_decomp(num_cls, label_vol):
    """lib.py:75-92: one-hot of an integer label map.  Accepts numpy (host, like the reference) or a torch
    tensor (stays on its device).  Returns float32 [..., num_cls]."""
    if isinstance(label_vol, np.ndarray):
        return np.stack([(label_vol == i) for i in range(num_cls)], axis=-1).astype(np.float32)
    return F.one_hot(label_vol, num_cls)


def _dice_eval(logits_or_pred, labels, n_class):
    """lib.py:96-110: hard Dice of argmax prediction vs one-hot labels, background included.
    Takes the logits (or the softmax map -- same argmax) [B,H,W,C]; returns (mean dice, per-class list)
    as device tensors computed from the confusion matrix kernel."""
    cm = F.confusion_counts(logits_or_pred, labels).to(torch.float64)
    inse = torch.diagonal(cm)
    union = cm.sum(0) + cm.sum(1)
    arr = 2.0 * inse / (union + 1e-7)
    return arr.mean(), [arr[i] for i in range(n_class)]


def _jaccard(conf_matrix):
    """lib.py:121-135"""
    cm = np.asarray(conf_matrix, dtype=np.float64)
    pp, gp, hit = cm.sum(0), cm.sum(1), np.diag(cm)
    den = pp + gp - hit
    return np.where(den == 0, 0.0, hit / np.where(den == 0, 1.0, den))


def _dice(conf_matrix):
    """lib.py:138-152"""
    cm = np.asarray(conf_matrix, dtype=np.float64)
    pp, gp, hit = cm.sum(0), cm.sum(1), np.diag(cm)
    den = pp + gp
    return np.where(den == 0, 0.0, 2.0 * hit / np.where(den == 0, 1.0, den))


def _indicator_eval(cm):
    """lib.py:155-175: print per-organ Dice / Jaccard from a confusion matrix."""
    contour_map = {"bg": 0, "la_myo": 1, "la_blood": 2, "lv_blood": 3, "aa": 4}
    dice, jaccard = _dice(cm), _jaccard(cm)
    print(cm)
    for organ, ind in contour_map.items():
        print("organ: %s" % organ)
        print("dice: %s" % dice[int(ind)])
        print("jaccard: %s" % jaccard[int(ind)])
    return dice, jaccard


def _save(state, model_path, global_step=None):
    """lib.py:23-29 analogue: variables keyed by TF names in one .npz (the checkpoint naming contract)."""
    import os
    path = model_path if global_step is None else "%s-%d" % (model_path, int(global_step))
    tmp = "%s.tmp.%d.npz" % (path, os.getpid())
    np.savez(tmp, **state)            # a reader (or a second writer) never sees a half-written checkpoint
    os.replace(tmp, path + ".npz")
    return path + ".npz"
