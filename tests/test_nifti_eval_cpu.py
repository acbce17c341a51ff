"""CPU tests of the test protocol's host side:

* the NIfTI-1 reader / writer behind lib.read_nii_image / write_nii / _save_nii_prediction (lib.py:31-72) -- byte-level known answers
  built here with `struct` from the NIfTI-1 header layout (independent of the writer under test), both byte orders, scaling, the
  three affine sources, gzip, error cases, round trips;
* `evaluation.run_test_eval` / `sample_metric_stddev` against tests/golden/reference_eval_vectors.json, which was produced by EXECUTING
  the reference's own `Trainer.test_eval` and `Trainer.sample_metric_stddev` (adversarial.py:993-1084) around a network-free stand-in
  predictor (tests/golden/make_reference_eval_vectors.py): frame lists, global-RNG shuffling, batch count, the zero rows of a short
  last batch, summed confusion matrix, per-subject Dice / Jaccard and the `[:1]` slip of the returned Jaccard value."""
import gzip
import json
import os
import struct
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _pkg():
    import pnp_b200  # noqa: F401
    from pnp_b200 import nifti, lib, evaluation
    return nifti, lib, evaluation


def _raw_header(bo, shape, code, bitpix, pixdim, vox_offset=352.0, slope=0.0, inter=0.0, qform=0, sform=0, quatern=(0, 0, 0),
                qoffset=(0, 0, 0), srow=None, magic=b"n+1\0"):
    """a NIfTI-1 header from the field offsets of nifti1.h (sizeof_hdr 0, dim 40, datatype 70, bitpix 72, pixdim 76, vox_offset 108,
    scl_slope 112, scl_inter 116, qform_code 252, sform_code 254, quatern_b 256, qoffset_x 268, srow_x 280, magic 344)"""
    h = bytearray(352)
    struct.pack_into(bo + "i", h, 0, 348)
    struct.pack_into(bo + "8h", h, 40, *([len(shape)] + list(shape) + [1] * (7 - len(shape))))
    struct.pack_into(bo + "2h", h, 70, code, bitpix)
    struct.pack_into(bo + "8f", h, 76, *pixdim)
    struct.pack_into(bo + "3f", h, 108, vox_offset, slope, inter)
    struct.pack_into(bo + "2h", h, 252, qform, sform)
    struct.pack_into(bo + "3f", h, 256, *quatern)
    struct.pack_into(bo + "3f", h, 268, *qoffset)
    if srow is not None:
        struct.pack_into(bo + "12f", h, 280, *np.asarray(srow, np.float64).reshape(-1))
    h[344:348] = magic
    return bytes(h)


@pytest.mark.parametrize("bo", ["<", ">"])
def test_nifti_reader_known_answer_both_byte_orders(tmp_path, bo):
    nifti, lib, _ = _pkg()
    shape = (3, 4, 2)
    vox = np.arange(24, dtype=np.int16) * 3 - 7                      # file order: x fastest
    srow = [[-1.5, 0, 0, 10], [0, 2.0, 0, -20], [0, 0, 2.5, 30]]
    blob = _raw_header(bo, shape, 4, 16, (1, 1.5, 2.0, 2.5, 1, 1, 1, 1), sform=1, srow=srow) + vox.astype(bo + "i2").tobytes()
    fid = str(tmp_path / "a.nii")
    open(fid, "wb").write(blob)
    img = lib.read_nii_image(fid)
    assert img.shape == shape and img.dtype == np.int16
    for x in range(3):
        for y in range(4):
            for z in range(2):
                assert img[x, y, z] == vox[x + 3 * y + 12 * z]
    obj = lib.read_nii_object(fid)
    exp = np.eye(4)
    exp[:3] = srow
    np.testing.assert_allclose(obj.get_affine(), exp)
    assert obj.header["byteorder"] == bo and obj.shape == shape


def test_nifti_scaling_gzip_and_offset(tmp_path):
    nifti, lib, _ = _pkg()
    vox = np.arange(8, dtype=np.uint8)
    pad = b"\0" * 16                                                 # vox_offset beyond the minimum: an extension block
    blob = _raw_header("<", (2, 2, 2), 2, 8, (1, 1, 1, 1, 1, 1, 1, 1), vox_offset=368.0, slope=0.5, inter=-1.0) + pad + vox.tobytes()
    fid = str(tmp_path / "s.nii.gz")
    with gzip.open(fid, "wb") as f:
        f.write(blob)
    img = lib.read_nii_image(fid)
    assert img.dtype == np.float64                                   # get_data() of a scaled image
    np.testing.assert_allclose(img.reshape(-1, order="F"), vox * 0.5 - 1.0)
    # slope 1 / inter 0 and slope 0 both mean "stored values": the dtype is kept
    for slope in (0.0, 1.0):
        blob = _raw_header("<", (2, 2, 2), 2, 8, (1,) * 8, slope=slope) + vox.tobytes()
        fid2 = str(tmp_path / ("u%d.nii" % int(slope)))
        open(fid2, "wb").write(blob)
        assert lib.read_nii_image(fid2).dtype == np.uint8


def test_nifti_affine_sources(tmp_path):
    nifti, lib, _ = _pkg()
    vox = np.zeros(24, np.float32)

    def affine(**kw):
        blob = _raw_header("<", (2, 3, 4), 16, 32, kw.pop("pixdim", (1, 2, 3, 4, 1, 1, 1, 1)), **kw) + vox.tobytes()
        fid = str(tmp_path / "q.nii")
        open(fid, "wb").write(blob)
        return lib.read_nii_object(fid).get_affine()

    # qform, zero quaternion: R = I scaled by the voxel sizes, offsets in the last column
    np.testing.assert_allclose(affine(qform=1, qoffset=(5, 6, 7)), [[2, 0, 0, 5], [0, 3, 0, 6], [0, 0, 4, 7], [0, 0, 0, 1]], atol=1e-6)
    # 90 degrees about z: (b, c, d) = (0, 0, sin 45) -> x axis maps to +y, y axis to -x
    s = np.sin(np.pi / 4)
    np.testing.assert_allclose(affine(qform=1, quatern=(0, 0, s)), [[0, -3, 0, 0], [2, 0, 0, 0], [0, 0, 4, 0], [0, 0, 0, 1]], atol=1e-6)
    # qfac = pixdim[0] = -1 flips the third axis
    np.testing.assert_allclose(affine(qform=1, pixdim=(-1, 2, 3, 4, 1, 1, 1, 1)), np.diag([2, 3, -4, 1]), atol=1e-6)
    # 180 degrees about x: (b, c, d) = (1, 0, 0), a = 0
    np.testing.assert_allclose(affine(qform=1, quatern=(1, 0, 0)), np.diag([2, -3, -4, 1]), atol=1e-6)
    # sform wins over qform
    srow = [[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3]]
    np.testing.assert_allclose(affine(qform=1, sform=2, srow=srow)[:3], srow)
    # neither: voxel sizes on the diagonal, x flipped, centred on the volume
    np.testing.assert_allclose(affine(), [[-2, 0, 0, 1.0], [0, 3, 0, -3.0], [0, 0, 4, -6.0], [0, 0, 0, 1]], atol=1e-6)


def test_nifti_rejects_what_it_cannot_read(tmp_path):
    nifti, lib, _ = _pkg()
    good = _raw_header("<", (2, 2, 2), 2, 8, (1,) * 8) + bytes(8)
    cases = {
        "short": good[:100],
        "magic": good[:344] + b"abc\0" + good[348:],
        "pair": _raw_header("<", (2, 2, 2), 2, 8, (1,) * 8, magic=b"ni1\0") + bytes(8),
        "sizeof": b"\1\2\3\4" + good[4:],
        "rgb": _raw_header("<", (2, 2, 2), 128, 24, (1,) * 8) + bytes(24),
        "truncated": good[:-3],
    }
    for name, blob in cases.items():
        fid = str(tmp_path / (name + ".nii"))
        open(fid, "wb").write(blob)
        with pytest.raises(ValueError):
            lib.read_nii_image(fid)
    with pytest.raises(IOError):
        lib.read_nii_image(str(tmp_path / "absent.nii"))


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.float32, np.float64, np.int64])
@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
def test_nifti_round_trip(tmp_path, dtype, ext):
    nifti, lib, _ = _pkg()
    rng = np.random.RandomState(3)
    a = (rng.standard_normal((5, 4, 3)) * 50).astype(dtype)
    aff = np.array([[0.8, 0.1, 0, -30], [-0.1, 0.9, 0, 12], [0, 0, 2.5, 7], [0, 0, 0, 1]])
    fid = lib.write_nii(a, "v" + ext, str(tmp_path), affine=aff)
    obj = lib.read_nii_object(fid)
    assert obj.get_data().dtype == np.dtype(dtype)
    np.testing.assert_array_equal(obj.get_data(), a)
    np.testing.assert_allclose(obj.get_affine(), aff, rtol=1e-6, atol=1e-6)
    raw = (gzip.open if ext.endswith(".gz") else open)(fid, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\0" and len(raw) == 352 + a.nbytes
    assert struct.unpack("<2h", raw[252:256]) == (0, 2)                # qform unset, sform "aligned": Nifti1Image(array, affine)


def test_write_nii_without_affine_and_save_prediction(tmp_path, capsys):
    nifti, lib, _ = _pkg()
    a = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    fid = lib.write_nii(a, "plain.nii", str(tmp_path))
    assert "No information about the global coordinate system" in capsys.readouterr().out
    np.testing.assert_allclose(lib.read_nii_object(fid).get_affine(), np.eye(4))
    with pytest.raises(Exception, match="cannot be saved"):
        lib.write_nii(a, "x.nii", str(tmp_path / "no" / "such" / "dir"))
    aff = np.diag([2.0, 2.0, 3.0, 1.0])
    ref = lib.write_nii(a.astype(np.float32), "ct_1003_image.nii.gz", str(tmp_path), affine=aff)
    gth = np.array([[[0, 1], [7, 4]], [[5, 2], [3, 9]]], np.int16)
    pred = np.array([[[0, 1], [2, 4]], [[0, 2], [3, 1]]], np.int16)
    p, g = lib._save_nii_prediction(gth, pred, ref, str(tmp_path), "dense_pred_ct_1003_image.nii.gz")
    assert os.path.basename(p) == "dense_pred_ct_1003_image.nii.gz" and os.path.basename(g) == "gth_dense_pred_ct_1003_image.nii.gz"
    np.testing.assert_array_equal(lib.read_nii_image(p), pred)
    np.testing.assert_array_equal(lib.read_nii_image(g), np.where(gth > 4, 0, gth))      # labels above the class range -> background
    np.testing.assert_allclose(lib.read_nii_object(g).get_affine(), aff)
    assert gth[0, 1, 0] == 7                                                               # the caller's array is not modified
    assert lib._inverse_lookup({"bg": 0, "aa": 4}, 4) == "aa" and lib._inverse_lookup({"bg": 0}, 3) is None


def test_subject_batches_known_answers():
    _, _, ev = _pkg()
    # depth 7, batch 2: frames 1..5, floor(7 / 2) = 3 batches, the last one short by one row
    assert ev.subject_batches(7, 2, False) == [[1, 2], [3, 4], [5]]
    # depth 9, batch 4: 7 usable frames, 2 batches of 4 and 3; depth 5, batch 3: one batch
    assert ev.subject_batches(9, 4, False) == [[1, 2, 3, 4], [5, 6, 7]]
    assert ev.subject_batches(5, 3, False) == [[1, 2, 3]]
    # depth 12, batch 5: 10 frames fill exactly floor(12 / 5) = 2 batches; depth 4, batch 4: one batch of 2 frames
    assert ev.subject_batches(12, 5, False) == [[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]]
    assert ev.subject_batches(4, 4, False) == [[1, 2]]
    assert ev.subject_batches(3, 4, False) == []
    rng = np.random.RandomState(9)
    got = ev.subject_batches(7, 2, True, rng)
    frames = list(range(1, 6))
    np.random.RandomState(9).shuffle(frames)
    assert got == [frames[0:2], frames[2:4], frames[4:6]]


GOLD = json.load(open(os.path.join(HERE, "golden", "reference_eval_vectors.json")))


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_test_eval_protocol_matches_executed_reference(tmp_path, name, capsys):
    """evaluation.run_test_eval + sample_metric_stddev on NIfTI files == the reference's own test_eval / sample_metric_stddev executed
    on the same seeded subjects with the same stand-in predictor"""
    import make_reference_eval_vectors as gen
    nifti, lib, ev = _pkg()
    case = GOLD["cases"][name]
    B, depths, flip, seed = case["batch_size"], case["depths"], case["flip_correction"], case["seed"]
    nii, lab = [], []
    for i, d in enumerate(depths):
        raw, raw_y = gen.make_subject(1000 * seed + i, d)
        nii.append(lib.write_nii(raw, "img_%d.nii.gz" % i, str(tmp_path)))
        lab.append(lib.write_nii(raw_y, "lab_%d.nii" % i, str(tmp_path)))
    calls = []

    def predict(vol, sl):
        assert vol.shape == (B, 256, 256, 3) and sl.shape == (B, 256, 256)
        pred = gen.stand_in_prediction(vol)
        cm = np.zeros((5, 5), np.int64)
        np.add.at(cm, (sl.reshape(-1), pred.reshape(-1)), 1)
        calls.append(int(np.count_nonzero(np.abs(vol).reshape(B, -1).sum(1))))
        return pred, cm

    out = str(tmp_path / "out")
    os.makedirs(out)
    np.random.seed(seed)                                   # the reference shuffles through the global numpy RNG
    sample_eval_list, all_cm = ev.run_test_eval(predict, lab, nii, B, 5, out, "dense_pred", flip_correction=flip, save_result=True,
                                                shuffle=True, write_cm=True)
    dice_list, jac_quirk = ev.sample_metric_stddev(sample_eval_list, 5)
    printed = capsys.readouterr().out
    assert calls == case["nonzero_rows_per_call"] and len(calls) == case["forward_calls"]
    np.testing.assert_array_equal(all_cm, np.asarray(case["all_cm"]))
    np.testing.assert_array_equal(np.loadtxt(os.path.join(out, "cm.csv")), np.asarray(case["all_cm"]))
    np.testing.assert_allclose(dice_list, case["subject_dice_list"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(jac_quirk, case["subject_jaccard_quirk"], rtol=0, atol=1e-15)
    assert np.asarray(jac_quirk).shape == (1, 2)
    d_mean, j_mean = ev.subject_level_means(sample_eval_list, 5)
    np.testing.assert_allclose(d_mean, dice_list)
    assert j_mean.shape == (5,) and abs(j_mean[0] - np.asarray(jac_quirk)[0, 1]) < 1e-15
    assert "all_dice_mean: " in printed and "organ: la_myo" in printed and "jaccard_stddev: " in printed
    # the saved dense predictions: evaluated orientation, reference image's affine, frames that were never fed stay 0
    for i, d in enumerate(depths):
        p = lib.read_nii_image(os.path.join(out, "dense_pred", "dense_pred_img_%d.nii.gz" % i))
        g = lib.read_nii_image(os.path.join(out, "dense_pred", "gth_dense_pred_img_%d.nii.gz" % i))
        raw, raw_y = gen.make_subject(1000 * seed + i, d)
        if flip:
            raw, raw_y = np.flip(np.flip(raw, 0), 1), np.flip(np.flip(raw_y, 0), 1)
        np.testing.assert_array_equal(g, raw_y)
        assert p.shape == raw_y.shape and not p[..., 0].any() and not p[..., d - 1].any()
        fed = [jj for jj in range(1, d - 1) if p[..., jj].any()]
        for jj in fed:
            np.testing.assert_array_equal(p[..., jj], gen.stand_in_prediction(raw[None, ..., jj - 1:jj + 2])[0])


def test_test_eval_argument_errors(tmp_path):
    _, lib, ev = _pkg()
    with pytest.raises(ValueError):
        ev.run_test_eval(lambda v, s: None, None, None, 2, 5, str(tmp_path), "dense_pred")
    with pytest.raises(Exception, match="cannot find sample"):
        ev.run_test_eval(lambda v, s: None, ["l.nii"], [str(tmp_path / "missing.nii")], 2, 5, str(tmp_path), "dense_pred")
    with pytest.raises(ValueError):
        ev.eval_volume(lambda v, s: None, np.zeros((4, 4, 3)), np.zeros((4, 4, 2)), 2, 5)
