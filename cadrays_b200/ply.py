"""PLY triangle-mesh reader / writer: the mesh format CADRays' exporter writes next to model.tcl
(`aiExportScene(..., "plyb", ...)`, src/ImportExport/AisMesh.cxx:490; `rtmeshread $Root/meshes/x.ply name`,
src/ImportExport/ImportExport.cxx:84-93).  ASCII and binary_little_endian; vertex properties x y z
[nx ny nz] [s t | u v | texture_u texture_v]; faces as index lists (polygons are fan-triangulated)."""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np

_TYPES = {"char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H", "uint16": "H",
          "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f", "double": "d", "float64": "d"}


def read_ply(path: str) -> Tuple[np.ndarray, Optional[np.ndarray], Optional[np.ndarray], np.ndarray]:
    """Returns (pos (n,3) f32, nrm (n,3) f32 or None, uv (n,2) f32 or None, idx (m,3) u32)."""
    data = open(path, "rb").read()
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    header = data[:end].decode("ascii", "replace").splitlines()
    if not header or header[0].strip() != "ply":
        raise ValueError("not a PLY file")
    fmt = None
    elements = []            # (name, count, [(kind, name, types...)])
    for line in header[1:]:
        t = line.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append((t[1], int(t[2]), []))
        elif t[0] == "property":
            if t[1] == "list":
                elements[-1][2].append(("list", t[4], t[2], t[3]))
            else:
                elements[-1][2].append(("scalar", t[2], t[1]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"unsupported PLY format {fmt}")
    E = ">" if fmt == "binary_big_endian" else "<"
    verts = None
    faces = []
    if fmt == "ascii":
        tokens = data[end:].split()
        pos_t = 0
        for name, count, props in elements:
            rows = []
            for _ in range(count):
                row = {}
                for p in props:
                    if p[0] == "scalar":
                        row[p[1]] = float(tokens[pos_t]); pos_t += 1
                    else:
                        n = int(tokens[pos_t]); pos_t += 1
                        row[p[1]] = [int(float(v)) for v in tokens[pos_t:pos_t + n]]; pos_t += n
                rows.append(row)
            if name == "vertex":
                verts = rows
            elif name == "face":
                faces = rows
        vcols = {k: np.array([r[k] for r in verts], dtype=np.float64) for k in (verts[0].keys() if verts else [])}
        face_lists = [next(iter(v for v in r.values() if isinstance(v, list))) for r in faces]
    else:
        off = end
        vcols, face_lists = {}, []
        fast_tris = None
        for name, count, props in elements:
            if all(p[0] == "scalar" for p in props):
                dt = np.dtype([(p[1], E + _TYPES[p[2]]) for p in props])
                arr = np.frombuffer(data, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                if name == "vertex":
                    vcols = {n: arr[n].astype(np.float64) for n in arr.dtype.names}
            else:
                # fast path: one list property and every face a triangle (what exporters write): fixed-size records
                if name == "face" and len(props) == 1 and count > 0:
                    cdt, idt = np.dtype(E + _TYPES[props[0][2]]), np.dtype(E + _TYPES[props[0][3]])
                    rec = np.dtype([("n", cdt), ("i", idt, (3,))])
                    if off + rec.itemsize * count <= len(data):
                        arr = np.frombuffer(data, dtype=rec, count=count, offset=off)
                        if (arr["n"] == 3).all():
                            fast_tris = arr["i"].astype(np.int64)
                            off += rec.itemsize * count
                            continue
                for _ in range(count):
                    row_list = None
                    for p in props:
                        if p[0] == "scalar":
                            off += struct.calcsize(E + _TYPES[p[2]])
                        else:
                            (n,) = struct.unpack_from(E + _TYPES[p[2]], data, off)
                            off += struct.calcsize(E + _TYPES[p[2]])
                            f = E + str(n) + _TYPES[p[3]]
                            row_list = list(struct.unpack_from(f, data, off))
                            off += struct.calcsize(f)
                    if name == "face" and row_list is not None:
                        face_lists.append(row_list)
    if not vcols or "x" not in vcols:
        raise ValueError("PLY has no vertex positions")
    pos = np.stack([vcols["x"], vcols["y"], vcols["z"]], 1).astype(np.float32)
    nrm = np.stack([vcols["nx"], vcols["ny"], vcols["nz"]], 1).astype(np.float32) if "nx" in vcols else None
    uv = None
    for a, b in (("s", "t"), ("u", "v"), ("texture_u", "texture_v")):
        if a in vcols and b in vcols:
            uv = np.stack([vcols[a], vcols[b]], 1).astype(np.float32)
            break
    tris = []
    for f in face_lists:
        for k in range(1, len(f) - 1):
            tris.append((f[0], f[k], f[k + 1]))
    idx = np.array(tris, dtype=np.int64).reshape(-1, 3)
    if fmt != "ascii" and fast_tris is not None:
        idx = fast_tris if idx.size == 0 else np.concatenate([fast_tris, idx])
    if idx.size and idx.min() < 0:
        raise ValueError("PLY face index out of range")
    idx = idx.astype(np.uint32)
    if idx.size and idx.max() >= pos.shape[0]:
        raise ValueError("PLY face index out of range")
    return pos, nrm, uv, idx


def write_ply(path: str, pos, nrm, idx, binary: bool = True, uv=None) -> None:
    pos = np.asarray(pos, np.float32); idx = np.asarray(idx, np.uint32)
    has_n = nrm is not None
    hdr = ["ply", "format " + ("binary_little_endian 1.0" if binary else "ascii 1.0"), f"element vertex {pos.shape[0]}",
           "property float x", "property float y", "property float z"]
    if has_n:
        hdr += ["property float nx", "property float ny", "property float nz"]
    if uv is not None:
        hdr += ["property float s", "property float t"]
    hdr += [f"element face {idx.shape[0]}", "property list uchar uint vertex_index", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        v = np.hstack([pos, np.asarray(nrm, np.float32)]) if has_n else pos
        if uv is not None:
            v = np.hstack([v, np.asarray(uv, np.float32)])
        if binary:
            f.write(v.astype("<f4").tobytes())
            rec = np.zeros(idx.shape[0], dtype=np.dtype([("n", "u1"), ("i", "<u4", (3,))]))
            rec["n"] = 3
            rec["i"] = idx
            f.write(rec.tobytes())
        else:
            for r in v:
                f.write((" ".join(repr(float(x)) for x in r) + "\n").encode())
            for t in idx:
                f.write(f"3 {t[0]} {t[1]} {t[2]}\n".encode())
