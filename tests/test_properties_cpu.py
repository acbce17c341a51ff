"""Property tests (hypothesis) of the host-side codecs and protocol helpers: NIfTI-1 round trips, TFRecord framing, event-file
encoding, the test protocol's frame batching, and the two oracle forms of 'SAME' pooling -- random shapes / dtypes / contents instead
of the hand-picked cases of the known-answer tests."""
import os

import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

FAST = settings(max_examples=30, deadline=None)


def _pkg():
    import pnp_b200  # noqa: F401
    from pnp_b200 import nifti, tfrecord, summary, evaluation
    return nifti, tfrecord, summary, evaluation


@FAST
@given(shape=st.lists(st.integers(1, 6), min_size=1, max_size=4), dtype=st.sampled_from(["u1", "i2", "i4", "f4", "f8", "i8", "u2", "i1"]),
       seed=st.integers(0, 2 ** 31 - 1), gz=st.booleans())
def test_nifti_round_trip_any_shape_and_dtype(tmp_path_factory, shape, dtype, seed, gz):
    nifti, _, _, _ = _pkg()
    rng = np.random.RandomState(seed)
    while len(shape) > 3 and shape[-1] == 1:                     # trailing singleton axes beyond 3-D are squeezed by the reader (documented)
        shape = shape[:-1]
    a = (rng.standard_normal(shape) * 100).astype(np.dtype(dtype))
    aff = np.eye(4)
    aff[:3, :] = rng.standard_normal((3, 4)).astype(np.float32)          # sform rows are stored as float32
    fid = str(tmp_path_factory.mktemp("nii") / ("v.nii.gz" if gz else "v.nii"))
    nifti.save(a, aff, fid)
    vol = nifti.load(fid)
    assert vol.get_data().dtype == a.dtype and vol.shape == a.shape
    np.testing.assert_array_equal(vol.get_data(), a)
    np.testing.assert_array_equal(vol.get_affine(), aff)
    assert vol.header["dim"][0] == a.ndim and vol.header["sform_code"] == 2 and vol.header["qform_code"] == 0


@FAST
@given(payloads=st.lists(st.binary(min_size=0, max_size=300), min_size=0, max_size=6))
def test_tfrecord_framing_round_trip(tmp_path_factory, payloads):
    _, tfr, _, _ = _pkg()
    p = str(tmp_path_factory.mktemp("rec") / "r.tfrecords")
    tfr.write_record(p, payloads)
    assert list(tfr.read_records(p)) == payloads
    assert os.path.getsize(p) == sum(16 + len(b) for b in payloads)            # 8 length + 4 crc + payload + 4 crc


@FAST
@given(items=st.lists(st.tuples(st.text(min_size=1, max_size=20), st.floats(allow_nan=False, width=32)), min_size=1, max_size=8),
       step=st.integers(-2 ** 40, 2 ** 40))
def test_event_encoding_parses_with_the_tensorboard_protos(items, step):
    pytest.importorskip("tensorboard")
    from tensorboard.compat.proto import event_pb2
    _, _, summary, _ = _pkg()
    ev = event_pb2.Event.FromString(summary.encode_event(123.5, step=step, summary=summary.encode_scalar_summary(items)))
    assert ev.wall_time == 123.5 and ev.step == step
    assert [(v.tag, v.simple_value) for v in ev.summary.value] == [(t, float(np.float32(x))) for t, x in items]


@FAST
@given(depth=st.integers(1, 40), batch=st.integers(1, 12), seed=st.integers(0, 1000))
def test_subject_batches_properties(depth, batch, seed):
    """adversarial.py:1021-1030: floor(depth / batch) batches cut from the (shuffled) frames 1 .. depth-2; no frame twice, no frame 0 or
    depth-1, every batch at most `batch` long, and every usable frame is fed exactly when depth - 2 <= floor(depth / batch) * batch"""
    _, _, _, ev = _pkg()
    for shuffle in (False, True):
        bs = ev.subject_batches(depth, batch, shuffle, np.random.RandomState(seed))
        flat = [f for b in bs for f in b]
        assert len(bs) == depth // batch and all(len(b) <= batch for b in bs)
        assert len(set(flat)) == len(flat) and all(1 <= f <= depth - 2 for f in flat)
        usable = max(depth - 2, 0)
        assert len(flat) == min(usable, (depth // batch) * batch)
        if not shuffle:
            assert flat == list(range(1, 1 + len(flat)))


@FAST
@given(H=st.integers(1, 12), W=st.integers(1, 12), n=st.integers(1, 6), avg=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_same_pooling_oracle_forms_agree(H, W, n, avg, seed):
    import torch
    from oracle import tf14_numpy as N, tf14_torch as T
    x = np.random.RandomState(seed).standard_normal((2, H, W, 3))
    a = N.pool_same(x, n, avg)
    b = T.pool_same(torch.from_numpy(x), n, avg).numpy()
    assert a.shape == b.shape == (2, -(-H // n), -(-W // n), 3)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    if n == 1:
        np.testing.assert_array_equal(a, x)
    # every input element belongs to exactly one window: max pooling never invents a value, the average stays inside the range
    assert a.max() <= x.max() + 1e-12 and a.min() >= x.min() - 1e-12
