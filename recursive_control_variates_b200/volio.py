"""Mitsuba `.vol` grid files (the format of the reference's scene volumes and of utils.save_final_state,
utils.py:45-60; written / read upstream by mi.VolumeGrid):

    bytes 0-2   'V' 'O' 'L'
    byte  3     version = 3
    int32       encoding = 1 (float32)
    int32 x 3   xres, yres, zres
    int32       channels
    float32 x 6 bounding box (xmin, ymin, zmin, xmax, ymax, zmax)
    float32[]   data, index ((z * yres + y) * xres + x) * channels + c   (little endian)

which is exactly the (Z, Y, X, C) layout the kernels and the plugins use."""
from __future__ import annotations

import struct

import numpy as np


def write_vol(path: str, data, bbox_min=(0.0, 0.0, 0.0), bbox_max=(1.0, 1.0, 1.0)) -> None:
    a = np.ascontiguousarray(np.asarray(data.detach().cpu() if hasattr(data, "detach") else data, dtype="<f4"))
    if a.ndim == 3:
        a = a[..., None]
    if a.ndim != 4:
        raise ValueError(f"expected a (Z,Y,X,C) grid, got shape {a.shape}")
    nz, ny, nx, c = a.shape
    with open(path, "wb") as f:
        f.write(b"VOL" + bytes([3]))
        f.write(struct.pack("<iiiii", 1, nx, ny, nz, c))
        f.write(struct.pack("<6f", *[float(v) for v in bbox_min], *[float(v) for v in bbox_max]))
        f.write(a.tobytes())


def read_vol(path: str):
    """Returns (data (Z,Y,X,C) float32, bbox_min, bbox_max)."""
    with open(path, "rb") as f:
        head = f.read(4)
        if head[:3] != b"VOL" or head[3] != 3:
            raise ValueError(f"{path}: not a version-3 Mitsuba volume file")
        enc, nx, ny, nz, c = struct.unpack("<iiiii", f.read(20))
        if enc != 1:
            raise ValueError(f"{path}: unsupported encoding {enc} (only float32 = 1)")
        if min(nx, ny, nz, c) <= 0:
            raise ValueError(f"{path}: invalid resolution {(nx, ny, nz, c)}")
        bbox = struct.unpack("<6f", f.read(24))
        n = nx * ny * nz * c
        raw = f.read(4 * n)
        if len(raw) != 4 * n:
            raise ValueError(f"{path}: truncated volume data ({len(raw)} of {4 * n} bytes)")
    data = np.frombuffer(raw, dtype="<f4").reshape(nz, ny, nx, c).astype(np.float32)
    return data, tuple(bbox[:3]), tuple(bbox[3:])
