"""Single-file NIfTI-1 volumes (.nii / .nii.gz) without nibabel -- what lib.py:47-72 of the reference needs from it.

The reference reads its test subjects with `nib.load(fid).get_data()` and writes predictions with `nib.Nifti1Image(array, affine)
.to_filename(fid)` (lib.py:31-72); `nibabel` is not installed here, and the format is a fixed 348-byte header followed by the voxels in
x-fastest (Fortran) order, so the reader / writer below cover exactly that: any of the scalar NIfTI datatypes, either byte order,
`scl_slope` / `scl_inter` scaling applied the way `get_data()` applies it, and the voxel-to-world affine chosen the way nibabel's
`get_affine()` chooses it (sform when `sform_code` > 0, else the qform quaternion when `qform_code` > 0, else the pixdim diagonal).
Header fields follow the NIfTI-1 specification (nifti1.h); offsets are bytes from the start of the file."""
import gzip
import os
import struct

import numpy as np

# NIfTI-1 datatype code -> numpy scalar type (codes of nifti1.h; RGB and complex-long-double are not scalar volumes)
_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 32: np.complex64, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}
HEADER_BYTES = 348


class Nifti1Volume(object):
    """What `nib.load()` returns, reduced to the three calls the reference makes: get_data(), get_affine(), .shape"""

    def __init__(self, data, affine, header):
        self._data, self._affine, self.header = data, affine, header

    def get_data(self):
        return self._data

    get_fdata = get_data

    def get_affine(self):
        return self._affine

    affine = property(get_affine)

    @property
    def shape(self):
        return self._data.shape


def _open(fid, mode):
    return gzip.open(fid, mode) if str(fid).endswith(".gz") else open(fid, mode)


def _quaternion_affine(b, c, d, offsets, pixdim):
    """qform of nifti1.h: rotation from the unit quaternion (a, b, c, d), a = sqrt(1 - b^2 - c^2 - d^2); pixdim[0] = qfac flips z"""
    a2 = 1.0 - (b * b + c * c + d * d)
    if a2 < 1e-7:                                   # 180 degree rotation: renormalise (b, c, d), a = 0
        s = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d, a = b * s, c * s, d * s, 0.0
    else:
        a = np.sqrt(a2)
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]], np.float64)
    qfac = -1.0 if pixdim[0] < 0 else 1.0
    M = np.eye(4)
    M[:3, :3] = R * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac], np.float64)[None, :]
    M[:3, 3] = offsets
    return M


def parse_header(raw):
    """348 header bytes -> dict (dim, datatype, pixdim, vox_offset, scl_slope, scl_inter, qform/sform, byte order '<' | '>')"""
    if len(raw) < HEADER_BYTES:
        raise ValueError("NIfTI-1 header needs %d bytes, file has %d" % (HEADER_BYTES, len(raw)))
    for bo in ("<", ">"):
        if struct.unpack(bo + "i", raw[0:4])[0] == HEADER_BYTES:
            break
    else:
        raise ValueError("not a NIfTI-1 file: sizeof_hdr is neither 348 nor byte-swapped 348")
    magic = bytes(raw[344:348])
    if magic not in (b"n+1\0", b"ni1\0"):
        raise ValueError("not a NIfTI-1 file: magic %r" % magic)
    if magic == b"ni1\0":
        raise ValueError("two-file NIfTI (.hdr/.img) is not supported; the reference's subjects are single .nii(.gz) files")
    h = {"byteorder": bo}
    h["dim"] = struct.unpack(bo + "8h", raw[40:56])
    h["datatype"], h["bitpix"] = struct.unpack(bo + "2h", raw[70:74])
    h["pixdim"] = struct.unpack(bo + "8f", raw[76:108])
    h["vox_offset"], h["scl_slope"], h["scl_inter"] = struct.unpack(bo + "3f", raw[108:120])
    h["qform_code"], h["sform_code"] = struct.unpack(bo + "2h", raw[252:256])
    h["quatern"] = struct.unpack(bo + "3f", raw[256:268])
    h["qoffset"] = struct.unpack(bo + "3f", raw[268:280])
    h["srow"] = np.array(struct.unpack(bo + "12f", raw[280:328]), np.float64).reshape(3, 4)
    nd = h["dim"][0]
    if not 1 <= nd <= 7:
        raise ValueError("NIfTI-1 dim[0] = %d out of range" % nd)
    if h["datatype"] not in _DTYPES:
        raise ValueError("NIfTI-1 datatype %d is not a scalar volume type" % h["datatype"])
    return h


def header_affine(h):
    """nibabel's get_best_affine(): sform, else qform, else the pixdim diagonal centred on the volume"""
    if h["sform_code"] > 0:
        M = np.eye(4)
        M[:3, :] = h["srow"]
        return M
    if h["qform_code"] > 0:
        return _quaternion_affine(*h["quatern"], h["qoffset"], h["pixdim"])
    nd = h["dim"][0]
    shape = np.array(list(h["dim"][1:1 + min(nd, 3)]) + [1] * (3 - min(nd, 3)), np.float64)
    zooms = np.array([abs(z) if z != 0 else 1.0 for z in h["pixdim"][1:4]], np.float64)
    M = np.diag(list(zooms) + [1.0])
    M[:3, 3] = -zooms * (shape - 1) / 2.0
    M[0, :] *= -1.0                                 # nibabel's fallback is radiological (x flipped)
    return M


def load(fid):
    """nib.load(fid): the whole file is read (the reference calls get_data() right away)"""
    if not os.path.isfile(fid):
        raise IOError("cannot find NIfTI file %s" % str(fid))
    with _open(fid, "rb") as f:
        blob = f.read()
    h = parse_header(blob[:HEADER_BYTES])
    nd = h["dim"][0]
    shape = tuple(int(s) for s in h["dim"][1:1 + nd])
    while len(shape) > 3 and shape[-1] == 1:        # trailing singleton time / vector axes: nibabel keeps them, the reference's
        shape = shape[:-1]                          # subjects are 3-D; squeezing only trailing ones keeps [256, 256, D] intact
    dt = np.dtype(_DTYPES[h["datatype"]]).newbyteorder(h["byteorder"])
    off = int(h["vox_offset"]) if h["vox_offset"] >= HEADER_BYTES else HEADER_BYTES + 4
    n = int(np.prod(shape, dtype=np.int64))
    if len(blob) < off + n * dt.itemsize:
        raise ValueError("NIfTI file %s is truncated: %d voxels of %s expected after byte %d" % (fid, n, dt, off))
    data = np.frombuffer(blob, dt, n, off).reshape(shape, order="F")
    slope, inter = h["scl_slope"], h["scl_inter"]
    if np.isfinite(slope) and slope != 0 and np.isfinite(inter) and not (slope == 1.0 and inter == 0.0):
        data = data.astype(np.float64) * slope + inter            # get_data() scales; unscaled files keep their stored dtype
    else:
        data = data.astype(dt.newbyteorder("="))
    return Nifti1Volume(data, header_affine(h), h)


def save(array_data, affine, fid):
    """nib.Nifti1Image(array, affine).to_filename(fid): sform = affine (code 2), qform unset, no scaling"""
    a = np.asarray(array_data)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    key = a.dtype.str[1:]
    if key not in _CODES:
        raise ValueError("dtype %s has no NIfTI-1 datatype code" % a.dtype)
    if not 1 <= a.ndim <= 7:
        raise ValueError("NIfTI-1 stores 1 to 7 dimensions, got %d" % a.ndim)
    affine = np.eye(4) if affine is None else np.asarray(affine, np.float64)
    if affine.shape != (4, 4):
        raise ValueError("affine must be 4 x 4")
    hdr = bytearray(352)                                          # 348 header bytes + the 4-byte empty extension flag
    struct.pack_into("<i", hdr, 0, HEADER_BYTES)
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<2h", hdr, 70, _CODES[key], a.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(0))
    pixdim = [1.0] + [float(z) for z in zooms] + [1.0] * 4
    struct.pack_into("<8f", hdr, 76, *pixdim)
    struct.pack_into("<3f", hdr, 108, 352.0, 0.0, 0.0)            # vox_offset; scl_slope = 0: "no scaling"
    hdr[123] = 2                                                  # xyzt_units: millimetres
    struct.pack_into("<2h", hdr, 252, 0, 2)                       # qform_code 0, sform_code 2 (aligned)
    struct.pack_into("<12f", hdr, 280, *[float(v) for v in affine[:3, :].reshape(-1)])
    hdr[344:348] = b"n+1\0"
    payload = a.astype(a.dtype.newbyteorder("<")).tobytes(order="F")
    with _open(fid, "wb") as f:
        f.write(bytes(hdr))
        f.write(payload)
    return fid
