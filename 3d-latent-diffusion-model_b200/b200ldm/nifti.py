"""Minimal NIfTI-1 (.nii) writer / reader: the egress format of the reference's inference CLI
(`nib.save(nib.Nifti1Image(volume[..., None], np.eye(4)), filename)`, 3d_ldm/inference.py:100-102).
nibabel is not installed here, and a single-file float32 volume needs 352 bytes of header."""
from __future__ import annotations

import struct

import numpy as np

_DT = {np.dtype("float32"): (16, 32), np.dtype("float64"): (64, 64), np.dtype("int16"): (4, 16),
       np.dtype("uint8"): (2, 8), np.dtype("int32"): (8, 32)}


def save_nifti(array, path: str, affine=None) -> str:
    """Write `array` (any rank <= 7, e.g. [D,H,W,1] as the reference does) with `affine` (default identity, as
    `np.eye(4)` in the reference) as a single-file little-endian NIfTI-1 image; data are stored first-index-fastest."""
    a = np.asarray(array)
    if a.dtype not in _DT:
        a = a.astype(np.float32)
    if a.ndim < 1 or a.ndim > 7:
        raise ValueError("NIfTI-1 supports 1..7 dimensions")
    aff = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    code, bits = _DT[a.dtype]
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    pixdim = [1.0] + [float(np.linalg.norm(aff[:3, i])) if i < 3 else 1.0 for i in range(7)]
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, code)
    struct.pack_into("<h", hdr, 72, bits)
    struct.pack_into("<8f", hdr, 76, *pixdim)
    struct.pack_into("<f", hdr, 108, 352.0)          # vox_offset
    struct.pack_into("<f", hdr, 112, 1.0)            # scl_slope
    struct.pack_into("<B", hdr, 123, 2)              # xyzt_units: mm
    struct.pack_into("<h", hdr, 252, 0)              # qform_code: unknown
    struct.pack_into("<h", hdr, 254, 2)              # sform_code: aligned
    struct.pack_into("<4f", hdr, 280, *aff[0])
    struct.pack_into("<4f", hdr, 296, *aff[1])
    struct.pack_into("<4f", hdr, 312, *aff[2])
    hdr[344:348] = b"n+1\x00"
    if not path.endswith((".nii",)):
        path = path + ".nii"
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\x00\x00\x00\x00")                 # no extensions
        f.write(np.asfortranarray(a).astype(a.dtype.newbyteorder("<")).tobytes(order="F"))
    return path


def load_nifti(path: str):
    """(array, affine) of a single-file little-endian NIfTI-1 image written by save_nifti (tests / round trips)."""
    raw = open(path, "rb").read()
    if struct.unpack_from("<i", raw, 0)[0] != 348 or raw[344:347] != b"n+1":
        raise ValueError("not a little-endian single-file NIfTI-1 image")
    dim = struct.unpack_from("<8h", raw, 40)
    code = struct.unpack_from("<h", raw, 70)[0]
    dt = {v[0]: k for k, v in _DT.items()}[code]
    off = int(struct.unpack_from("<f", raw, 108)[0])
    shape = tuple(dim[1:1 + dim[0]])
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt.newbyteorder("<"), count=n, offset=off).reshape(shape, order="F")
    aff = np.eye(4)
    aff[0], aff[1], aff[2] = (struct.unpack_from("<4f", raw, o) for o in (280, 296, 312))
    return np.array(data), aff
