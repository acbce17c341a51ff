"""TensorBoard event files and the TFRecord framing against an INDEPENDENT implementation: the `tensorboard` package that ships in this
image (its own event-file loader, protobuf classes and `RecordWriter` / `masked_crc32c`; no TensorFlow needed).

* `summary.FileWriter` (tf.summary.FileWriter + add_summary of the merged scalar op, adversarial.py:664-697,989-991) writes files
  that tensorboard's loader parses: version event first, then one event per monitoring pass carrying the reference's tags in order;
* records framed by tensorboard's RecordWriter are read by the native reader (libpnp_io.so: pnp_tfrecord_count / decode) and the
  masked CRC32C of both implementations agrees -- the TFRecord side of SURVEY 8(f) row 2 pinned to TF-lineage code."""
import os
import struct

import numpy as np
import pytest

tb = pytest.importorskip("tensorboard")


def _events(path):
    from tensorboard.backend.event_processing.event_file_loader import RawEventFileLoader
    from tensorboard.compat.proto import event_pb2
    return [event_pb2.Event.FromString(raw) for raw in RawEventFileLoader(path).Load()]


def test_event_file_is_read_by_tensorboard(tmp_path):
    import pnp_b200  # noqa: F401
    from pnp_b200 import summary
    from pnp_b200.adversarial import Trainer
    w = summary.FileWriter(str(tmp_path / "train_loggan-0.3_x"))
    tags = Trainer.SCALAR_TAGS
    assert tags == ("fixed_coeff_reg", "discriminator_loss", "generator_loss", "ct_dice_eval_c1_lv_myo", "ct_dice_eval_c2_la_blood",
                    "ct_dice_eval_c3_lv_blood", "ct_dice_eval_c4_aa", "mri_dice", "learning_rate")        # adversarial.py:665-692
    rows = []
    for step in (0, 5, 300, 2 ** 40 + 7):
        vals = [float(np.float32(np.sin(step + i) * 10 ** (i - 4))) for i in range(len(tags))]
        rows.append((step, vals))
        w.add_scalars(dict(zip(tags, vals)), step)
    w.add_summary(summary.encode_scalar_summary([("extra", -1.5)]), global_step=-3)       # serialized-Summary form, negative step
    w.close()
    files = os.listdir(w.logdir)
    assert len(files) == 1 and files[0].startswith("events.out.tfevents.")
    evs = _events(w.path)
    assert len(evs) == 6
    assert evs[0].file_version == "brain.Event:2" and evs[0].wall_time > 1.6e9
    for ev, (step, vals) in zip(evs[1:5], rows):
        assert ev.step == step and ev.WhichOneof("what") == "summary"
        assert [v.tag for v in ev.summary.value] == list(tags)
        assert [v.simple_value for v in ev.summary.value] == vals                      # float32 round trip is exact
    assert evs[5].step == -3 and evs[5].summary.value[0].tag == "extra" and evs[5].summary.value[0].simple_value == -1.5
    # the accumulator TensorBoard's UI is built on sees the scalars too
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(w.logdir)
    acc.Reload()
    assert set(acc.Tags()["scalars"]) >= set(tags)
    got = acc.Scalars("discriminator_loss")
    assert [g.step for g in got][:3] == [0, 5, 300] and got[1].value == pytest.approx(rows[1][1][1])


def test_segmenter_scalar_tags_follow_the_reference():
    import pnp_b200  # noqa: F401
    from pnp_b200.source_segmenter import Trainer
    assert Trainer.SCALAR_TAGS == ("loss", "regularizer_loss", "weighted_loss", "dice_loss", "dice_eval", "dice_eval_c1", "dice_eval_c2",
                                   "dice_eval_c3", "dice_eval_c4")                                        # source_segmenter.py:387-396


def test_masked_crc32c_and_framing_agree_with_tensorboard(tmp_path):
    import pnp_b200  # noqa: F401
    from pnp_b200 import tfrecord as R, _io
    from tensorboard.summary.writer.record_writer import RecordWriter, masked_crc32c
    rng = np.random.RandomState(7)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000, 4097):
        blob = rng.bytes(n)
        assert R.masked_crc(blob) == masked_crc32c(blob), n
    # records written by tensorboard's RecordWriter: our framing is byte-identical, and the native reader accepts theirs
    img = rng.standard_normal((256, 256, 3)).astype(np.float32)
    lab = rng.randint(0, 5, size=(256, 256, 3)).astype(np.float32)
    payloads = [R.encode_example(img, lab), R.encode_example(img[::-1].copy(), lab[:, ::-1].copy())]
    theirs, ours = str(tmp_path / "theirs.tfrecords"), str(tmp_path / "ours.tfrecords")
    with open(theirs, "wb") as f:
        rw = RecordWriter(f)
        for p in payloads:
            rw.write(p)
        rw.flush()
    R.write_record(ours, payloads)
    assert open(theirs, "rb").read() == open(ours, "rb").read()
    buf = open(theirs, "rb").read()
    assert _io.lib.pnp_tfrecord_count(buf, len(buf)) == 2
    out_img, out_lab = R.load_slice(theirs, record_index=1, check_crc=True)
    np.testing.assert_array_equal(np.asarray(out_img), img[::-1])
    np.testing.assert_array_equal(np.asarray(out_lab), lab[:, ::-1][:, :, 1].astype(np.int64))
    # a flipped payload bit is caught by the data CRC, a flipped length bit by the length CRC
    bad = bytearray(buf)
    bad[12 + 100] ^= 1
    p = str(tmp_path / "bad.tfrecords")
    open(p, "wb").write(bytes(bad))
    with pytest.raises(IOError):
        R.load_slice(p, record_index=0, check_crc=True)
    bad = bytearray(buf)
    bad[0] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(IOError):
        R.load_slice(p, record_index=0, check_crc=True)
    assert struct.unpack("<Q", buf[:8])[0] == len(payloads[0])


def test_event_bytes_known_answer():
    """protobuf wire format by hand: Event{wall_time = 1.0 (field 1, fixed64), step = 3 (field 2, varint), summary (field 5, bytes)
    {value (field 1) {tag = "a" (field 1), simple_value = 2.0 (field 2, fixed32)}}} and the TFRecord frame around it"""
    import pnp_b200  # noqa: F401
    from pnp_b200 import summary, tfrecord
    ev = summary.encode_event(1.0, step=3, summary=summary.encode_scalar_summary([("a", 2.0)]))
    assert ev == bytes.fromhex("09" "000000000000f03f" "10" "03" "2a" "0a" "0a" "08" "0a" "01" "61" "15" "00000040")
    fr = summary.frame(ev)
    assert fr[:8] == struct.pack("<Q", len(ev)) and fr[12:12 + len(ev)] == ev and len(fr) == len(ev) + 16
    assert struct.unpack("<I", fr[8:12])[0] == tfrecord.masked_crc(fr[:8]) and struct.unpack("<I", fr[-4:])[0] == tfrecord.masked_crc(ev)
    # masked CRC32C of the empty string: crc32c("") = 0 -> ((0 >> 15) | (0 << 17)) + 0xa282ead8
    assert tfrecord.masked_crc(b"") == 0xA282EAD8 and tfrecord.crc32c(b"123456789") == 0xE3069283          # RFC 3720 check value
    # negative steps are two's-complement 10-byte varints, version events carry field 3
    assert summary.encode_event(0.0, step=-1)[9:] == bytes.fromhex("10" + "ff" * 9 + "01")
    assert summary.encode_event(0.0, file_version="brain.Event:2")[9:] == b"\x1a\x0dbrain.Event:2"
