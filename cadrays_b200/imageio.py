"""Image writers for the dump path: PNG for BufferDump(Graphic3d_BT_RGB) and
Radiance .hdr / PFM for the float dump (AppGui.cxx:339-350,430,503).  Inputs are bottom-up rows, as
BufferDump returns them; files are written top-down."""
from __future__ import annotations

import struct
import zlib

import numpy as np


def write_png(path: str, rgb8_bottom_up: np.ndarray) -> None:
    img = np.ascontiguousarray(rgb8_bottom_up[::-1], dtype=np.uint8)
    h, w, c = img.shape
    assert c == 3
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def read_png_rgb8(path: str) -> np.ndarray:
    """Minimal reader for files written by write_png (8-bit RGB, filter 0); returns top-down rows."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = zlib.decompress(idat)
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 3).copy()


def write_hdr(path: str, rgb32f_bottom_up: np.ndarray) -> None:
    """Radiance RGBE, uncompressed scanlines."""
    img = np.ascontiguousarray(rgb32f_bottom_up[::-1], dtype=np.float32)
    h, w, _ = img.shape
    m = np.max(img, axis=2)
    e = np.zeros_like(m, dtype=np.int32)
    nz = m > 1e-32
    e[nz] = np.floor(np.log2(m[nz])).astype(np.int32) + 1
    scale = np.where(nz, np.ldexp(1.0, 8 - e), 0.0)[..., None]
    rgbe = np.zeros((h, w, 4), dtype=np.uint8)
    rgbe[..., :3] = np.clip(img * scale, 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(nz, e + 128, 0).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())


def write_pfm(path: str, rgb32f_bottom_up: np.ndarray) -> None:
    img = np.ascontiguousarray(rgb32f_bottom_up, dtype="<f4")   # PFM is bottom-up already
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(f"PF\n{w} {h}\n-1.0\n".encode())
        f.write(img.tobytes())
