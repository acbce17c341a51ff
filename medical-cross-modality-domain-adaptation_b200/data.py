"""Synthetic stand-in for the reference's TFRecord queues (source_segmenter.py:331-355,
adversarial.py:607-631): 256x256x3 z-scored slices + a 256x256 integer label map per sample
(README.md:49-64 schema).  The real MMWHS records are not available offline;
the TFRecord reader itself is a "next" row.

Batches are produced in pinned host memory, like a feed_dict, so the host->device copy is part of the
step exactly as in the reference (source_segmenter.py:484).
"""
import numpy as np
import torch


def label_maps(B, seed, size=256, num_cls=5):
    """nested ellipses: classes 1..num_cls-1 on background 0, centres / radii jittered by `seed`"""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.zeros((B, size, size), np.int64)
    radii = np.linspace(0.39 * size, 0.1 * size, num_cls - 1)
    for b in range(B):
        cy, cx = size / 2 + rng.uniform(-20, 20), size / 2 + rng.uniform(-20, 20)
        for c, rad in zip(range(1, num_cls), radii):
            ry, rx = rad * rng.uniform(0.8, 1.0), rad * rng.uniform(0.8, 1.0)
            out[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = c
    return out


class SyntheticSource:
    """yields (images [B,256,256,3] fp32, labels [B,256,256] int64) as pinned host tensors"""

    def __init__(self, batch_size, seed=1234, shift=0.0, scale=1.0, size=256, num_cls=5, pool=4, contrast=0.0):
        """contrast > 0 adds a per-class intensity ((class - (num_cls-1)/2) * contrast / 2) to the noise, so that the labels can be
        learned from the images (the held-out Dice gate trains on this); 0 = pure z-scored noise (the throughput workloads)"""
        self.B, self.size, self.num_cls = batch_size, size, num_cls
        g = torch.Generator().manual_seed(seed)
        self.pool = []
        for i in range(pool):
            x = torch.randn(batch_size, size, size, 3, generator=g) * scale + shift
            y = torch.from_numpy(label_maps(batch_size, seed + 99 + i, size, num_cls))
            if contrast:
                x = x + ((y.to(torch.float32) - (num_cls - 1) / 2.0) * (contrast / 2.0)).unsqueeze(-1)
            if torch.cuda.is_available():
                x, y = x.pin_memory(), y.pin_memory()
            self.pool.append((x, y))
        self.i = 0

    def next(self):
        item = self.pool[self.i % len(self.pool)]
        self.i += 1
        return item


def to_device(x, dev):
    return x.to(dev, non_blocking=True)
