"""TFRecord input for the reference's data format (README.md:49-64; parser at source_segmenter.py:331-355 /
adversarial.py:607-631), without TensorFlow: the record framing (length, masked CRC32C, payload, masked CRC32C) and the
tf.train.Example protobuf wire format are decoded by hand.

Each example carries 8 features: dsize_dim0/1/2, lsize_dim0/1/2 (int64) and data_vol / label_vol (raw float32 bytes of a
256x256x3 volume).  Like the reference, the image is the full 3-slice volume and the label is the MIDDLE slice
(`tf.slice(label_vol, [0,0,1], [256,256,1])`, source_segmenter.py:348); the reference reads lsize_dim2 from dsize_dim2
(:341) -- harmless because both are 3.
"""
import struct

import numpy as np
import torch

# ---- CRC32C (Castagnoli), table driven -----------------------------------------------------------------------------
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


def crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


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


class TFRecordSource:
    """Drop-in for data.SyntheticSource: shuffled batches (images [B,256,256,3] fp32, labels [B,256,256] int64) in pinned
    host memory from a list of single-example TFRecord files (lists/*_list), like tf.train.shuffle_batch."""

    def __init__(self, file_list, batch_size, seed=0):
        self.files = list(file_list)
        self.B = batch_size
        self.rng = np.random.RandomState(seed)
        self.order = self.rng.permutation(len(self.files))
        self.pos = 0

    def _next_file(self):
        if self.pos >= len(self.order):
            self.order = self.rng.permutation(len(self.files))
            self.pos = 0
        f = self.files[self.order[self.pos]]
        self.pos += 1
        return f

    def next(self):
        xs, ys = [], []
        while len(xs) < self.B:
            for payload in read_records(self._next_file()):
                x, y = decode_slice(payload)
                xs.append(x)
                ys.append(y)
                if len(xs) == self.B:
                    break
        x = torch.from_numpy(np.stack(xs))
        y = torch.from_numpy(np.stack(ys))
        if torch.cuda.is_available():
            x, y = x.pin_memory(), y.pin_memory()
        return x, y


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
