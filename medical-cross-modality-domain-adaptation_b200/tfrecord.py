"""TFRecord input for the reference's data format (README.md:49-64; parser at source_segmenter.py:331-355 /
adversarial.py:607-631), without TensorFlow: the record framing (length, masked CRC32C, payload, masked CRC32C) and the
tf.train.Example protobuf wire format are decoded by hand.

Each example carries 8 features: dsize_dim0/1/2, lsize_dim0/1/2 (int64) and data_vol / label_vol (raw float32 bytes of a
256x256x3 volume).  Like the reference, the image is the full 3-slice volume and the label is the MIDDLE slice
(`tf.slice(label_vol, [0,0,1], [256,256,1])`, source_segmenter.py:348); the reference reads lsize_dim2 from dsize_dim2
(:341) -- harmless because both are 3.
"""
import struct
import threading

import numpy as np
import torch

from . import _io


# ---- CRC32C (Castagnoli): libpnp_io.so (SSE4.2 crc32 instruction, slicing-by-8 fallback) -----------------------------
def crc32c(data):
    data = bytes(data)
    return int(_io.lib.pnp_crc32c(data, len(data)))


def masked_crc(data):
    data = bytes(data)
    return int(_io.lib.pnp_masked_crc32c(data, len(data)))


# ---- protobuf wire helpers --------------------------------------------------------------------------------------------
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """yield (field_number, wire_type, value) of one message; length-delimited values come back as memoryview slices"""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def parse_example(payload):
    """tf.train.Example -> {name: bytes | [int] | [float]}.  Example{1: Features{1: map<string, Feature>}};
    Feature{1: BytesList{1: bytes*}, 2: FloatList{1: packed float}, 3: Int64List{1: packed varint}}"""
    out = {}
    buf = memoryview(payload)
    for fn, wt, feats in _fields(buf):
        if fn != 1:
            continue
        for fn2, wt2, entry in _fields(feats):
            if fn2 != 1:
                continue
            key, feature = None, None
            for fn3, wt3, v in _fields(entry):
                if fn3 == 1:
                    key = bytes(v).decode()
                elif fn3 == 2:
                    feature = v
            val = None
            for kind, wt4, lst in _fields(feature):
                if kind == 1:      # BytesList
                    items = [bytes(v) for f, w, v in _fields(lst) if f == 1]
                    val = items[0] if len(items) == 1 else items
                elif kind == 3:    # Int64List (packed or not)
                    ints = []
                    for f, w, v in _fields(lst):
                        if f != 1:
                            continue
                        if w == 0:
                            ints.append(v)
                        else:
                            p = 0
                            while p < len(v):
                                x, p = _varint(v, p)
                                ints.append(x)
                    val = ints
                elif kind == 2:    # FloatList
                    fl = []
                    for f, w, v in _fields(lst):
                        if f == 1:
                            fl += list(np.frombuffer(bytes(v), dtype="<f4"))
                    val = fl
            out[key] = val
    return out


def read_records(path, check_crc=True):
    """iterate the raw payloads of one TFRecord file"""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if len(head) < 12:
                return
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if check_crc and masked_crc(head[:8]) != lcrc:
                raise IOError("%s: corrupt record length" % path)
            data = f.read(length)
            (dcrc,) = struct.unpack("<I", f.read(4))
            if check_crc and masked_crc(data) != dcrc:
                raise IOError("%s: corrupt record payload" % path)
            yield data


def decode_slice(payload, raw_size=(256, 256, 3)):
    """one example -> (image float32 [256,256,3], label int64 [256,256]) with the reference's slicing"""
    ex = parse_example(payload)
    vol = np.frombuffer(ex["data_vol"], dtype="<f4").reshape(raw_size)
    lab = np.frombuffer(ex["label_vol"], dtype="<f4").reshape(raw_size)
    return vol.copy(), lab[:, :, 1].astype(np.int64)


def load_slice(path, record_index=0, check_crc=True, raw_size=(256, 256, 3), image=None, label=None):
    """one example of a TFRecord file -> (image float32 [H,W,C], label int64 [H,W]) decoded natively
    (pnp_tfrecord_load_file: framing + CRC + protobuf + decode_raw + the reference's slicing), into the given arrays"""
    H, W, C = raw_size
    if image is None:
        image = np.empty(raw_size, np.float32)
    if label is None:
        label = np.empty((H, W), np.int64)
    rc = _io.lib.pnp_tfrecord_load_file(path.encode(), record_index, 1 if check_crc else 0, image.ctypes.data, label.ctypes.data,
                                        H, W, C, 1)
    _io.check(rc, path)
    return image, label


class TFRecordSource:
    """Drop-in for data.SyntheticSource: shuffled batches (images [B,256,256,3] fp32, labels [B,256,256] int64) in pinned
    host memory from a list of single-example TFRecord files (lists/*_list), like the reference's input pipeline
    (source_segmenter.py:331-355): `num_threads` reader threads (reference: 4 QueueRunner threads) walk a shuffled file order
    and decode natively with the GIL released into a shuffle buffer of `capacity` examples; next() draws a batch at random once
    more than `min_after_dequeue` examples would remain (tf.train.shuffle_batch(capacity=120, min_after_dequeue=30)) and hands
    it out in one of two alternating pinned buffers, so the previous batch's host->device copy may still be in flight.
    num_threads=0: synchronous, deterministic order (tests)."""

    def __init__(self, file_list, batch_size, seed=0, num_threads=4, capacity=120, min_after_dequeue=30, check_crc=True,
                 raw_size=(256, 256, 3)):
        self.files = list(file_list)
        if not self.files:
            raise ValueError("empty TFRecord file list")
        self.B = batch_size
        self.raw_size = tuple(raw_size)
        self.check_crc = check_crc
        self.rng = np.random.RandomState(seed)
        self.order = self.rng.permutation(len(self.files))
        self.pos = 0
        self.capacity = max(capacity, min_after_dequeue + batch_size)
        self.min_after = min_after_dequeue
        H, W, C = self.raw_size
        pin = torch.cuda.is_available()
        self._out = [(torch.empty(batch_size, H, W, C, dtype=torch.float32, pin_memory=pin),
                      torch.empty(batch_size, H, W, dtype=torch.int64, pin_memory=pin)) for _ in range(2)]
        self._flip = 0
        self.examples_read = 0
        self._threads = []
        self._stop = False
        self._error = None
        if num_threads > 0:
            self._lock = threading.Lock()
            self._cv = threading.Condition(self._lock)
            self._free = [(np.empty(self.raw_size, np.float32), np.empty((H, W), np.int64)) for _ in range(self.capacity)]
            self._ready = []
            for _ in range(num_threads):
                t = threading.Thread(target=self._worker, daemon=True)
                t.start()
                self._threads.append(t)

    def _next_file(self):
        if self.pos >= len(self.order):
            self.order = self.rng.permutation(len(self.files))
            self.pos = 0
        f = self.files[self.order[self.pos]]
        self.pos += 1
        return f

    def _worker(self):
        try:
            while True:
                with self._cv:
                    while not self._free and not self._stop:
                        self._cv.wait()
                    if self._stop:
                        return
                    slot = self._free.pop()
                    path = self._next_file()
                load_slice(path, 0, self.check_crc, self.raw_size, slot[0], slot[1])      # native, GIL released
                with self._cv:
                    self._ready.append(slot)
                    self.examples_read += 1
                    self._cv.notify_all()
        except Exception as e:      # noqa: BLE001 -- surface reader errors in next() instead of dying silently
            with self._cv:
                self._error = e
                self._cv.notify_all()

    def next(self):
        x, y = self._out[self._flip]
        self._flip ^= 1
        if not self._threads:
            for i in range(self.B):
                load_slice(self._next_file(), 0, self.check_crc, self.raw_size, x[i].numpy(), y[i].numpy())
                self.examples_read += 1
            return x, y
        need = min(self.min_after + self.B, self.capacity)
        with self._cv:
            while len(self._ready) < need and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error
            picks = []
            for _ in range(self.B):
                j = int(self.rng.randint(len(self._ready)))
                self._ready[j], self._ready[-1] = self._ready[-1], self._ready[j]
                picks.append(self._ready.pop())
        xn, yn = x.numpy(), y.numpy()
        for i, (im, lb) in enumerate(picks):
            xn[i] = im
            yn[i] = lb
        with self._cv:
            self._free.extend(picks)
            self._cv.notify_all()
        return x, y

    def close(self):
        if self._threads:
            with self._cv:
                self._stop = True
                self._cv.notify_all()
            for t in self._threads:
                t.join(timeout=2)
            self._threads = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- writer (tests / synthetic dataset export) ---------------------------------------------------------------------
def _enc_varint(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(image, label_vol):
    """inverse of decode_slice's parser: the 8-feature schema of README.md:49-64"""
    feats = b""

    def add(name, feature):
        nonlocal feats
        feats += _ld(1, _ld(1, name.encode()) + _ld(2, feature))
    for i, d in enumerate(image.shape):
        add("dsize_dim%d" % i, _ld(3, _ld(1, _enc_varint(int(d)))))
    for i, d in enumerate(label_vol.shape):
        add("lsize_dim%d" % i, _ld(3, _ld(1, _enc_varint(int(d)))))
    add("data_vol", _ld(1, _ld(1, np.asarray(image, "<f4").tobytes())))
    add("label_vol", _ld(1, _ld(1, np.asarray(label_vol, "<f4").tobytes())))
    return _ld(1, feats)


def write_record(path, payloads):
    with open(path, "wb") as f:
        for p in payloads:
            head = struct.pack("<Q", len(p))
            f.write(head + struct.pack("<I", masked_crc(head)) + p + struct.pack("<I", masked_crc(p)))
