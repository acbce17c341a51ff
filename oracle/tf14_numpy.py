"""
ORACLE (test infrastructure, NOT product code) -- naive numpy / pure-loop restatement of the
TensorFlow-1.4 op semantics that the reference graph relies on.

PARITY UNPINNED: the reference (carrenD/Medical-Cross-Modality-Domain-Adaptation) ships no tests,
fixtures or golden vectors, and its runtime (tensorflow-gpu==1.4.0 / py2.7, README.md:20-25) cannot
be executed in this image.  These functions restate the published TF-1.4 behaviour of the ops the
reference calls (call sites cited per function) and are pinned only by hand-derived known-answer
tests (tests/test_oracle_kat.py) and by cross-checking against the independent torch-CPU form in
oracle/tf14_torch.py.

PINNED TO EXECUTED REFERENCE CODE (tests/golden/, generator scripts committed; tests/test_reference_golden.py,
tests/test_reference_graph_trace.py): phase_shift_literal / PS_literal / PS_closed_form and label_decomp bit-exactly
against the reference's own ops.py:3-27 and lib.py:75-92; softmax_weighted_loss and dice_loss against
source_segmenter.py:241-273 evaluated numerically.  Everything that is defined by TensorFlow kernels (conv padding
offsets, batch norm, dropout scaling, Adam / RMSProp) remains UNPINNED restatement.

CROSS-CHECKED AGAINST THIRD-PARTY TF SEMANTICS (tests/test_tf_semantics_opencv_cpu.py): conv2d ('SAME', every stride / kernel parity
class of the graphs), the dilated convolution against the SpaceToBatchND -> Conv2D -> BatchToSpaceND graph TF 1.4 emits for
atrous_conv2d, pool_same, inference-mode batch_norm, leaky_relu / relu, softmax -- executed by OpenCV's TensorFlow importer
(cv2.dnn.readNetFromTensorflow) on GraphDefs built from the TF protos in `tensorboard`.  Train-mode batch norm, dropout scaling and
the optimizers have no such counterpart in this image and stay unpinned.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product path (medical-cross-modality-domain-adaptation_b200/) never does.

Everything here is float64, NHWC activations, HWIO weights, loops written for clarity -- use only
on tiny shapes.
"""
import math
import numpy as np


# ----------------------------------------------------------------------------------------------
# padding rules
# ----------------------------------------------------------------------------------------------
def same_pad(in_size, k, stride, dilation=1):
    """TF 'SAME' padding amounts (before, after) along one spatial axis.
    Used by tf.nn.conv2d (layers.py:18,67) and tf.nn.atrous_conv2d (layers.py:86).
    out = ceil(in/s); total = max((out-1)*s + (k-1)*d + 1 - in, 0); before = total//2."""
    out = -(-in_size // stride)
    eff = (k - 1) * dilation + 1
    total = max((out - 1) * stride + eff - in_size, 0)
    return total // 2, total - total // 2


def symmetric_pad_index(i, n):
    """Index into an axis of length n for padded coordinate i (may be <0 or >=n), tf.pad(...,
    'SYMMETRIC') semantics = mirror INCLUDING the edge element (layers.py:23,72,91)."""
    if i < 0:
        return -i - 1
    if i >= n:
        return 2 * n - 1 - i
    return i


def symmetric_pad(x, p):
    """x: [B,H,W,C]; pad p rows/cols on every spatial side, edge-inclusive mirror."""
    B, H, W, C = x.shape
    out = np.zeros((B, H + 2 * p, W + 2 * p, C), dtype=x.dtype)
    for i in range(H + 2 * p):
        si = symmetric_pad_index(i - p, H)
        for j in range(W + 2 * p):
            sj = symmetric_pad_index(j - p, W)
            out[:, i, j, :] = x[:, si, sj, :]
    return out


# ----------------------------------------------------------------------------------------------
# convolution (direct loops)
# ----------------------------------------------------------------------------------------------
def conv2d(x, w, stride=1, dilation=1, padding="SAME"):
    """tf.nn.conv2d / tf.nn.atrous_conv2d restated (layers.py:18-24, 67-73, 86-92).
    x [B,H,W,Cin] NHWC, w [kh,kw,Cin,Cout] HWIO (cross-correlation, no kernel flip).
    padding: 'SAME' (TF asymmetric zero pad), 'VALID', or 'SYMMETRIC' (mirror pad k//2 then VALID,
    exactly what layers.py does before calling conv2d with padding='VALID')."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    kh, kw, cin, cout = w.shape
    if padding == "SYMMETRIC":
        assert kh == kw
        x = symmetric_pad(x, kh // 2)
        padding = "VALID"
    B, H, W, C = x.shape
    assert C == cin
    if padding == "SAME":
        pt, _ = same_pad(H, kh, stride, dilation)
        pl, _ = same_pad(W, kw, stride, dilation)
        Ho = -(-H // stride)
        Wo = -(-W // stride)
    elif padding == "VALID":
        pt = pl = 0
        Ho = (H - ((kh - 1) * dilation + 1)) // stride + 1
        Wo = (W - ((kw - 1) * dilation + 1)) // stride + 1
    else:
        raise ValueError(padding)
    y = np.zeros((B, Ho, Wo, cout), dtype=np.float64)
    for oy in range(Ho):
        for ox in range(Wo):
            acc = np.zeros((B, cout), dtype=np.float64)
            for ky in range(kh):
                iy = oy * stride + ky * dilation - pt
                if iy < 0 or iy >= H:
                    continue
                for kx in range(kw):
                    ix = ox * stride + kx * dilation - pl
                    if ix < 0 or ix >= W:
                        continue
                    acc += x[:, iy, ix, :] @ w[ky, kx]
            y[:, oy, ox, :] = acc
    return y


# ----------------------------------------------------------------------------------------------
# batch norm, activations, pooling
# ----------------------------------------------------------------------------------------------
BN_DECAY = 0.90   # layers.py:100
BN_EPS = 1e-3     # tf.contrib.layers.batch_norm default epsilon


def batch_norm(x, gamma, beta, moving_mean, moving_var, is_training):
    """tf.contrib.layers.batch_norm(decay=.9, scale, center, updates_collections=None) (layers.py:95-100).
    Returns (y, new_moving_mean, new_moving_var).  Train: normalise with the biased batch variance,
    feed the UNBIASED variance into the moving average (fused kernel behaviour)."""
    x = np.asarray(x, dtype=np.float64)
    C = x.shape[-1]
    flat = x.reshape(-1, C)
    n = flat.shape[0]
    if is_training:
        mean = flat.mean(axis=0)
        var = ((flat - mean) ** 2).mean(axis=0)
        y = (x - mean) / np.sqrt(var + BN_EPS) * gamma + beta
        unbiased = var * n / max(n - 1, 1)
        mm = BN_DECAY * moving_mean + (1 - BN_DECAY) * mean
        mv = BN_DECAY * moving_var + (1 - BN_DECAY) * unbiased
        return y, mm, mv
    y = (x - moving_mean) / np.sqrt(moving_var + BN_EPS) * gamma + beta
    return y, moving_mean, moving_var


def leaky_relu(x, alpha=0.2):
    """tf.nn.leaky_relu default alpha=0.2 (layers.py:12)."""
    return np.where(x > 0, x, alpha * x)


def relu(x):
    return np.maximum(x, 0)


def max_pool2x2(x):
    """tf.nn.max_pool ksize 2 stride 2 SAME on even dims (layers.py:102-103)."""
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def pool_same(x, n, avg=False):
    """tf.nn.max_pool / tf.nn.avg_pool with ksize = strides = [1,n,n,1], padding 'SAME' (layers.py:102-106), naive loops.
    Published TF behaviour (GetWindowedOutputSizeVerbose): out = ceil(in / n), pad_needed = out * n - in, pad_before = pad_needed // 2;
    padded positions never win a max and AvgPool divides by the number of VALID elements of a window."""
    B, H, W, C = x.shape
    Ho, Wo = -(-H // n), -(-W // n)
    pt, pl = (Ho * n - H) // 2, (Wo * n - W) // 2
    y = np.zeros((B, Ho, Wo, C), x.dtype)
    for oy in range(Ho):
        y0, y1 = max(oy * n - pt, 0), min(oy * n - pt + n, H)
        for ox in range(Wo):
            x0, x1 = max(ox * n - pl, 0), min(ox * n - pl + n, W)
            win = x[:, y0:y1, x0:x1, :].reshape(B, -1, C)
            y[:, oy, ox, :] = win.mean(1) if avg else win.max(1)
    return y


def channel_pad_skip(x):
    """tf.pad(x, [[0,0],[0,0],[0,0],[C//2, C//2]]) (layers.py:160,182)."""
    C = x.shape[-1]
    return np.pad(x, [(0, 0), (0, 0), (0, 0), (C // 2, C // 2)])


# ----------------------------------------------------------------------------------------------
# phase shift (ops.py) -- literal emulation of the TF op sequence in numpy
# ----------------------------------------------------------------------------------------------
def _tf_squeeze(a):
    return np.squeeze(a)  # tf.squeeze with no axis removes every size-1 dim, like numpy


def phase_shift_literal(I, r, batch_size):
    """Literal numpy emulation of ops.py:3-21 (_phase_shift) -- reshape/transpose/split/squeeze/
    concat executed exactly as written, including the batch_size==1 special cases."""
    _, a, b, c = I.shape
    X = np.reshape(I, (batch_size, a, b, r, r))
    X = np.transpose(X, (0, 1, 2, 4, 3))
    X = np.split(X, a, 1)
    X = np.concatenate([_tf_squeeze(x) for x in X], 2)
    if batch_size == 1:
        X = np.expand_dims(X, 0)
    X = np.split(X, b, 1)
    if batch_size == 1:
        X = np.concatenate([x for x in X], 2)
    else:
        X = np.concatenate([_tf_squeeze(x) for x in X], 2)
    out = np.reshape(X, (batch_size, a * r, b * r, 1))
    if batch_size == 1:
        out = np.transpose(out, (0, 2, 1, 3))
    return out


def PS_literal(X, r, n_channel, batch_size):
    """ops.py:23-27."""
    Xc = np.split(X, n_channel, -1)
    return np.concatenate([phase_shift_literal(x, r, batch_size) for x in Xc], 3)


def PS_closed_form(X, r, n_channel, batch_size):
    """Index law derived from the literal emulation:
       B>=2: out[n, i*r+q, j*r+p, g] = X[n, i, j, g*r*r + p*r + q]
       B==1: out[n, i*r+p, j*r+q, g] = X[n, i, j, g*r*r + p*r + q]"""
    B, a, b, C = X.shape
    assert B == batch_size and C == n_channel * r * r
    out = np.zeros((B, a * r, b * r, n_channel), dtype=X.dtype)
    for g in range(n_channel):
        for p in range(r):
            for q in range(r):
                src = X[:, :, :, g * r * r + p * r + q]
                if B >= 2:
                    out[:, q::r, p::r, g] = src
                else:
                    out[:, p::r, q::r, g] = src
    return out


# ----------------------------------------------------------------------------------------------
# losses / metrics (source_segmenter.py:241-273, lib.py:75-110)
# ----------------------------------------------------------------------------------------------
def softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def pixel_wise_softmax_2(x):
    """layers.py:134-138: exp/sum without max subtraction, clipped to +-1e15."""
    e = np.exp(x)
    return np.clip(e / e.sum(axis=3, keepdims=True), -1e15, 1e15)


def softmax_weighted_loss(logits, y):
    """source_segmenter.py:241-258."""
    p = softmax(np.asarray(logits, np.float64))
    n_class = y.shape[-1]
    raw = 0.0
    tot = y.sum()
    for i in range(n_class):
        gti = y[..., i]
        wi = 1 - gti.sum() / tot
        raw = raw + -1.0 * wi * gti * np.log(np.clip(p[..., i], 0.005, 1))
    return raw.mean()


def dice_loss(logits, y):
    """source_segmenter.py:260-273."""
    p = softmax(np.asarray(logits, np.float64))
    n_class = y.shape[-1]
    dice = 0.0
    for i in range(n_class):
        inse = (p[..., i] * y[..., i]).sum()
        l = (p[..., i] * p[..., i]).sum()
        r = y[..., i].sum()
        dice += 2.0 * inse / (l + r + 1e-7)
    return -dice / n_class


def label_decomp(num_cls, label_vol):
    """lib.py:75-92 numpy one-hot."""
    out = np.zeros(label_vol.shape + (num_cls,), dtype=np.float32)
    for i in range(num_cls):
        out[..., i] = (label_vol == i)
    return out


def dice_eval(compact_pred, labels, n_class):
    """lib.py:96-110."""
    pred = label_decomp(n_class, compact_pred).astype(np.float64)
    arr = []
    for i in range(n_class):
        inse = (pred[..., i] * labels[..., i]).sum()
        union = pred[..., i].sum() + labels[..., i].sum()
        arr.append(2.0 * inse / (union + 1e-7))
    return sum(arr) / n_class, arr


def l2_loss(w):
    """tf.nn.l2_loss = sum(w^2)/2 (source_segmenter.py:237)."""
    return float((np.asarray(w, np.float64) ** 2).sum() / 2)


# ----------------------------------------------------------------------------------------------
# optimizers (scalar/array forms)
# ----------------------------------------------------------------------------------------------
def adam_step(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (source_segmenter.py:378): epsilon-hat form.
    t is the 1-based step count AFTER increment."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def rmsprop_step(theta, g, ms, mom, lr=3e-4, decay=0.9, momentum=0.0, eps=1e-10):
    """tf.train.RMSPropOptimizer update (adversarial.py:643-652): ms starts at ONE, eps inside sqrt."""
    ms = decay * ms + (1 - decay) * g * g
    mom = momentum * mom + lr * g / np.sqrt(ms + eps)
    theta = theta - mom
    return theta, ms, mom


def truncated_normal(rng, shape, stddev):
    """tf.truncated_normal: N(0, stddev) with |z|>2 redrawn (layers.py:48,55)."""
    n = int(np.prod(shape))
    out = rng.standard_normal(n)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2
    return (out * stddev).reshape(shape).astype(np.float32)
