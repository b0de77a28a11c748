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
    """Reader for 8-bit, non-interlaced RGB / RGBA / grey PNG files (all five scanline filters): what write_png
    produces and what `rttexture` / the material icons of the reference use.  Returns top-down rows, RGB."""
    data = open(path, "rb").read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, w, h, depth, ctype, interlace = 8, b"", 0, 0, 8, 2, 0
    palette = None
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body[:13])
        elif tag == b"PLTE":
            palette = np.frombuffer(body, dtype=np.uint8).reshape(-1, 3)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or interlace != 0 or channels is None:
        raise ValueError(f"{path}: only 8-bit non-interlaced PNG files are supported")
    stride = w * channels
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + stride)
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros(stride, dtype=np.int32)
            for x in range(stride):
                a = cur[x - channels] if x >= channels else 0
                b = prev[x]
                c = prev[x - channels] if x >= channels else 0
                if f == 1:
                    pred = a
                elif f == 3:
                    pred = (a + b) >> 1
                elif f == 4:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError(f"{path}: bad scanline filter {f}")
                cur[x] = (line[x] + pred) & 255
        out[y] = cur
        prev = cur
    px = out.reshape(h, w, channels)
    if ctype == 3:
        if palette is None:
            raise ValueError(f"{path}: palette image without PLTE")
        return palette[px[..., 0]].copy()
    if channels == 1 or channels == 2:
        return np.repeat(px[..., :1], 3, axis=2).copy()
    return px[..., :3].copy()


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


def read_hdr(path: str) -> np.ndarray:
    """Radiance RGBE (.hdr / .pic), flat or new-style run-length encoded scanlines, -Y +X orientation.
    Returns float32 RGB, top-down rows (what vtextureenv hands to crt_envmap_set_rgb32f)."""
    data = open(path, "rb").read()
    if not data.startswith(b"#?"):
        raise ValueError(f"{path}: not a Radiance picture")
    end = data.index(b"\n\n") + 2
    line_end = data.index(b"\n", end)
    res = data[end:line_end].split()
    if len(res) != 4 or res[0] != b"-Y" or res[2] != b"+X":
        raise ValueError(f"{path}: unsupported orientation {data[end:line_end]!r}")
    h, w = int(res[1]), int(res[3])
    pos = line_end + 1
    rgbe = np.zeros((h, w, 4), dtype=np.uint8)
    for y in range(h):
        if 8 <= w < 32768 and data[pos] == 2 and data[pos + 1] == 2 and not (data[pos + 2] & 0x80):
            if (data[pos + 2] << 8 | data[pos + 3]) != w:
                raise ValueError(f"{path}: scanline width mismatch")
            pos += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = data[pos]; pos += 1
                    if n > 128:
                        n -= 128
                        rgbe[y, x:x + n, c] = data[pos]; pos += 1
                    else:
                        rgbe[y, x:x + n, c] = np.frombuffer(data, np.uint8, n, pos); pos += n
                    x += n
                if x != w:
                    raise ValueError(f"{path}: corrupt run-length data")
        else:
            rgbe[y] = np.frombuffer(data, np.uint8, 4 * w, pos).reshape(w, 4); pos += 4 * w
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * scale[..., None]).astype(np.float32)


def read_pfm(path: str) -> np.ndarray:
    """Portable float map (PF colour / Pf grey), either byte order; returns float32 RGB, top-down rows."""
    with open(path, "rb") as f:
        kind = f.readline().strip()
        dims = f.readline().split()
        while len(dims) < 2:
            dims += f.readline().split()
        w, h = int(dims[0]), int(dims[1])
        scale = float(f.readline().strip())
        ch = {b"PF": 3, b"Pf": 1}.get(kind)
        if ch is None:
            raise ValueError(f"{path}: not a PFM file")
        a = np.frombuffer(f.read(4 * w * h * ch), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, ch)
    img = a[::-1].astype(np.float32)          # PFM rows are bottom-up
    return np.repeat(img, 3, axis=2) if ch == 1 else img

