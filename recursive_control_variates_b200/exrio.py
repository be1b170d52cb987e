"""Minimal OpenEXR files (what utils.save_img / save_final_state write through mi.Bitmap upstream, utils.py:8-14,53-54):
single-part scanline images, 32-bit float channels, no compression — readable by any EXR viewer.

Layout (OpenEXR file format v2): magic 0x01312f76, version 2, attribute list (name, type, size, value) terminated by a
null byte, one 64-bit offset per scanline, then per scanline: y (int32), byte count (int32), the pixels channel after
channel in ALPHABETICAL channel order (B, G, R for an RGB image)."""
from __future__ import annotations

import struct

import numpy as np

_MAGIC = 20000630


def _attr(name: str, typ: str, payload: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path: str, img, channels: str = None) -> None:
    """img: (H, W) or (H, W, C) float array (numpy or torch); channels: one letter per channel (default Y / RGB / RGBA)."""
    a = np.asarray(img.detach().cpu() if hasattr(img, "detach") else img, dtype=np.float32)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    names = list(channels) if channels else {1: ["Y"], 3: ["R", "G", "B"], 4: ["R", "G", "B", "A"]}[c]
    if len(names) != c:
        raise ValueError("one channel name per channel")
    order = sorted(range(c), key=lambda i: names[i])                      # files store channels alphabetically
    chlist = b"".join(names[i].encode() + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for i in order) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    head = struct.pack("<ii", _MAGIC, 2)
    head += _attr("channels", "chlist", chlist)
    head += _attr("compression", "compression", b"\0")
    head += _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
    head += _attr("lineOrder", "lineOrder", b"\0")
    head += _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    head += _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    head += _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    head += b"\0"
    row_bytes = 4 * w * c
    first = len(head) + 8 * h
    offsets = struct.pack(f"<{h}Q", *[first + y * (8 + row_bytes) for y in range(h)])
    planar = np.ascontiguousarray(a[:, :, order].transpose(0, 2, 1)).astype("<f4")     # (H, C, W): channel after channel
    with open(path, "wb") as f:
        f.write(head); f.write(offsets)
        for y in range(h):
            f.write(struct.pack("<ii", y, row_bytes)); f.write(planar[y].tobytes())


def read_exr(path: str):
    """Reads what write_exr writes (uncompressed float32 scanlines); returns (image (H, W, C), channel names in file order)."""
    with open(path, "rb") as f:
        data = f.read()
    magic, version = struct.unpack_from("<ii", data, 0)
    if magic != _MAGIC or (version & 0xFF) != 2:
        raise ValueError(f"{path}: not an OpenEXR 2 file")
    pos, attrs = 8, {}
    while data[pos] != 0:
        e = data.index(b"\0", pos); name = data[pos:e].decode(); pos = e + 1
        e = data.index(b"\0", pos); typ = data[pos:e].decode(); pos = e + 1
        size, = struct.unpack_from("<i", data, pos); pos += 4
        attrs[name] = (typ, data[pos:pos + size]); pos += size
    pos += 1
    if attrs["compression"][1] != b"\0":
        raise ValueError(f"{path}: compressed EXR files are not supported")
    names, p, ch = [], 0, attrs["channels"][1]
    while ch[p] != 0:
        e = ch.index(b"\0", p); names.append(ch[p:e].decode())
        ptype, = struct.unpack_from("<i", ch, e + 1)
        if ptype != 2:
            raise ValueError(f"{path}: only 32-bit float channels are supported")
        p = e + 1 + 16
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h, c = x1 - x0 + 1, y1 - y0 + 1, len(names)
    offsets = struct.unpack_from(f"<{h}Q", data, pos)
    out = np.zeros((h, c, w), np.float32)
    for off in offsets:
        y, nbytes = struct.unpack_from("<ii", data, off)
        out[y - y0] = np.frombuffer(data, "<f4", c * w, off + 8).reshape(c, w)
    return out.transpose(0, 2, 1).copy(), names
