"""ctypes binding of include/pnp_io.h (libpnp_io.so: host-side TFRecord decoding in plain C, built in-tree by _build.py).
ctypes releases the GIL around every call, so the reader threads of tfrecord.py decode in parallel."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpnp_io.so")

if not os.path.exists(LIB_PATH):
    raise ImportError("libpnp_io.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
lib = ctypes.CDLL(LIB_PATH)

_P, _sz, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
SIGNATURES = {
    "pnp_crc32c": ([_P, _sz], ctypes.c_uint32),
    "pnp_crc32c_sw": ([_P, _sz], ctypes.c_uint32),
    "pnp_masked_crc32c": ([_P, _sz], ctypes.c_uint32),
    "pnp_crc32c_is_hardware": ([], _int),
    "pnp_tfrecord_count": ([_P, _sz], _int),
    "pnp_tfrecord_decode": ([_P, _sz, _int, _int, _P, _P, _int, _int, _int, _int], _int),
    "pnp_tfrecord_load_file": ([ctypes.c_char_p, _int, _int, _P, _P, _int, _int, _int, _int], _int),
    "pnp_io_error_string": ([_int], ctypes.c_char_p),
}
for _n, (_a, _r) in SIGNATURES.items():
    _f = getattr(lib, _n)
    _f.argtypes, _f.restype = _a, _r


class DecodeError(IOError):
    pass


def check(rc, what):
    if rc != 0:
        raise DecodeError("%s: [%d] %s" % (what, rc, lib.pnp_io_error_string(rc).decode()))
