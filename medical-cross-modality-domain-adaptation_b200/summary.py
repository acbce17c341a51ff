"""TensorBoard event files without TensorFlow: what `tf.summary.FileWriter(dir)` + `add_summary(scalar_summary_op, step)` leave on
disk (adversarial.py:664-697, 807-808, 989-991; source_segmenter.py:394-407, 452-453, 537-539).

An event file is a TFRecord stream (the framing `tfrecord.write_record` already implements: u64 length, masked CRC32C of the length,
payload, masked CRC32C of the payload -- CRC32C from libpnp_io.so) of serialized `tensorflow.Event` messages.  Only four fields are
needed and they are encoded by hand from event.proto / summary.proto:

    Event   { double wall_time = 1; int64 step = 2; oneof what { string file_version = 3; Summary summary = 5; } }
    Summary { repeated Value value = 1; }      Value { string tag = 1; float simple_value = 2; }

The first record of a file is the version event `brain.Event:2`, as TensorFlow writes it.  Image summaries (`tf.summary.image`) are
not written (out of scope).  tests/test_summary_cpu.py reads the files back with the `tensorboard` package's own loader."""
import os
import socket
import struct
import time

from .tfrecord import _enc_varint, _ld, masked_crc


def _key(field, wire):
    return _enc_varint((field << 3) | wire)


def encode_scalar_summary(scalars):
    """Summary message with one Value per (tag, number) pair, in the given order (the order of the merged summary op)"""
    out = b""
    for tag, v in scalars:
        value = _ld(1, str(tag).encode("utf-8")) + _key(2, 5) + struct.pack("<f", float(v))
        out += _ld(1, value)
    return out


def encode_event(wall_time, step=None, file_version=None, summary=None):
    ev = _key(1, 1) + struct.pack("<d", float(wall_time))
    if step is not None:
        ev += _key(2, 0) + _enc_varint(int(step) & 0xFFFFFFFFFFFFFFFF)
    if file_version is not None:
        ev += _ld(3, file_version.encode("utf-8"))
    if summary is not None:
        ev += _ld(5, summary)
    return ev


def frame(payload):
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", masked_crc(head)) + payload + struct.pack("<I", masked_crc(payload))


class FileWriter(object):
    """tf.summary.FileWriter for scalars: `events.out.tfevents.<seconds>.<hostname>` under `logdir`, appended and flushed per call."""

    def __init__(self, logdir, filename_suffix=""):
        os.makedirs(logdir, exist_ok=True)
        self.logdir = logdir
        now = time.time()
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s.%d%s" % (int(now), socket.gethostname(), os.getpid(),
                                                                              filename_suffix))
        self._f = open(self.path, "ab")
        self._f.write(frame(encode_event(now, file_version="brain.Event:2")))
        self._f.flush()

    def add_scalars(self, scalars, step, wall_time=None):
        """one Event carrying every (tag, value) pair: what add_summary(sess.run(tf.summary.merge(scalar_summaries)), step) writes.
        `scalars`: dict (insertion order kept) or a list of pairs."""
        items = list(scalars.items()) if hasattr(scalars, "items") else list(scalars)
        ev = encode_event(time.time() if wall_time is None else wall_time, step=step, summary=encode_scalar_summary(items))
        self._f.write(frame(ev))

    def add_summary(self, summary, global_step=None):
        """the reference's call shape: `summary` is a serialized Summary message (bytes) or a dict of scalars"""
        if isinstance(summary, (bytes, bytearray)):
            self._f.write(frame(encode_event(time.time(), step=global_step, summary=bytes(summary))))
        else:
            self.add_scalars(summary, global_step)

    def flush(self):
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.flush()
            self._f.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
