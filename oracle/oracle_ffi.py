"""ctypes wrapper of the CPU ORACLE (oracle/cadrays_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(cadrays_b200/) never imports this module.

oracle/_ref: the reference's own implementation of this path is OCCT (external,
unpinned, absent from /root/reference) -- nothing under /root/reference compiles
into a renderer, so there is no oracle/_ref build; PARITY IS UNPINNED.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REPO = ORACLE_DIR.parent
LIB_PATH = ORACLE_DIR / "_build" / "libcadrays_oracle.so"

sys.path.insert(0, str(REPO))
from cadrays_b200._ffi import crt_bsdf, crt_camera, crt_light, crt_params, crt_stats  # noqa: E402  (POD types only)

CFLAGS = ["-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall", "-Wextra"]


def build_oracle(force: bool = False) -> Path:
    src = ORACLE_DIR / "cadrays_oracle.c"
    hdr = ORACLE_DIR / "cadrays_oracle.h"
    if not force and LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= max(src.stat().st_mtime, hdr.stat().st_mtime):
        return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    cmd = [os.environ.get("ORACLE_CC", "gcc"), *CFLAGS, "-o", str(LIB_PATH), str(src), "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None
_f = C.POINTER(C.c_float)
_i32 = C.POINTER(C.c_int32)
_u8 = C.POINTER(C.c_uint8)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build_oracle()
        L = C.CDLL(str(LIB_PATH))
        L.orc_scene_from_blob.restype = C.c_void_p
        L.orc_scene_from_blob.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_set_materials.argtypes = [C.c_void_p, C.POINTER(crt_bsdf), C.c_uint32]
        L.orc_set_lights.argtypes = [C.c_void_p, C.POINTER(crt_light), C.c_uint32]
        L.orc_set_envmap_rgb32f.argtypes = [C.c_void_p, _f, C.c_uint32, C.c_uint32]
        L.orc_set_envmap_rgb8.argtypes = [C.c_void_p, _u8, C.c_uint32, C.c_uint32]
        L.orc_set_textures.argtypes = [C.c_void_p, _u8, C.POINTER(C.c_uint32), C.c_uint32]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(crt_params)]
        L.orc_set_camera.argtypes = [C.c_void_p, C.POINTER(crt_camera)]
        L.orc_trace.argtypes = [C.c_void_p, _f, _f, _f, C.c_uint32, C.c_int, _i32, _i32, _f, _f, _f, C.POINTER(crt_stats)]
        L.orc_trace_brute.argtypes = [C.c_void_p, _f, _f, _f, C.c_uint32, C.c_int, _i32, _i32, _f, _f, _f]
        L.orc_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, _f, C.c_int, C.POINTER(crt_stats)]
        _u32 = C.POINTER(C.c_uint32)
        L.orc_adaptive_allocate.argtypes = [_u32, C.c_uint32, C.c_uint32, C.c_uint32, _u32]
        L.orc_render_adaptive.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, _f,
                                          _u32, _u32, _f, _u32, C.c_int]
        L.orc_display.argtypes = [C.c_void_p, _f, C.c_uint32, C.c_uint32, _u8]
        L.orc_hdr.argtypes = [_f, C.c_uint32, C.c_uint32, _f]
        L.orc_sincos2pi.argtypes = [C.c_float, _f, _f]
        L.orc_exp.restype = C.c_float; L.orc_exp.argtypes = [C.c_float]
        L.orc_atan2.restype = C.c_float; L.orc_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_acos.restype = C.c_float; L.orc_acos.argtypes = [C.c_float]
        L.orc_bullard_frame_seed.restype = C.c_uint32; L.orc_bullard_frame_seed.argtypes = [C.c_uint32, C.c_uint64]
        L.orc_seed_rand.restype = C.c_uint32
        L.orc_seed_rand.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_rand_float.restype = C.c_float; L.orc_rand_float.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_fresnel.argtypes = [C.c_float, _f, _f]
        L.orc_bsdf_eval.argtypes = [C.POINTER(crt_bsdf), _f, _f, C.c_int, _f]
        L.orc_bsdf_pdf.restype = C.c_float; L.orc_bsdf_pdf.argtypes = [C.POINTER(crt_bsdf), _f, _f, _f]
        L.orc_bsdf_sample.restype = C.c_float
        L.orc_bsdf_sample.argtypes = [C.POINTER(crt_bsdf), _f, _f, _f, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.c_int]
        L.orc_camera_ray.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, _f, _f]
        L.orc_scene_epsilon.restype = C.c_float; L.orc_scene_epsilon.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f)


class OracleScene:
    """Oracle-side scene: the blob exported by the product's host builder (so both sides walk the
    same BVH bytes) plus the same material / light / camera / parameter records."""

    def __init__(self, blob: bytes):
        self._L = lib()
        self._blob = blob
        self._h = self._L.orc_scene_from_blob(blob, len(blob))
        if not self._h:
            raise ValueError("oracle rejected the blob")

    def close(self):
        if self._h:
            self._L.orc_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, desc):
        """Takes a cadrays_b200.scenes.SceneDesc (plain data) and sets the non-geometry state."""
        mats = (crt_bsdf * max(len(desc.materials), 1))()
        for i, b in enumerate(desc.materials):
            mats[i] = b.to_c()
        self._L.orc_set_materials(self._h, mats, len(desc.materials))
        ls = (crt_light * max(len(desc.lights), 1))()
        for i, l in enumerate(desc.lights):
            ls[i] = l
        self._L.orc_set_lights(self._h, ls, len(desc.lights))
        if desc.envmap is None:
            self._L.orc_set_envmap_rgb32f(self._h, None, 0, 0)
        elif desc.envmap.dtype == np.uint8:
            a = np.ascontiguousarray(desc.envmap[..., :3])
            self._L.orc_set_envmap_rgb8(self._h, a.ctypes.data_as(_u8), a.shape[1], a.shape[0])
        else:
            a = np.ascontiguousarray(desc.envmap[..., :3], dtype=np.float32)
            self._L.orc_set_envmap_rgb32f(self._h, _fp(a), a.shape[1], a.shape[0])
        texs = []
        for t in getattr(desc, "textures", []):
            t = np.ascontiguousarray(t, dtype=np.uint8)
            if t.shape[2] == 3:
                t = np.concatenate([t, np.full(t.shape[:2] + (1,), 255, np.uint8)], axis=2)
            texs.append(np.ascontiguousarray(t))
        if texs:
            blob = np.concatenate([t.reshape(-1) for t in texs])
            sizes = np.array([[t.shape[1], t.shape[0]] for t in texs], dtype=np.uint32).reshape(-1)
            self._L.orc_set_textures(self._h, blob.ctypes.data_as(_u8), sizes.ctypes.data_as(C.POINTER(C.c_uint32)), len(texs))
        else:
            self._L.orc_set_textures(self._h, None, None, 0)
        p = desc.params.to_c()
        self._L.orc_set_params(self._h, C.byref(p))
        desc.camera.Aspect = desc.width / desc.height
        c = desc.camera.to_c()
        self._L.orc_set_camera(self._h, C.byref(c))

    def set_params(self, params):
        p = params.to_c()
        self._L.orc_set_params(self._h, C.byref(p))

    def trace(self, org, dir, tmax=None, any_hit=False, brute=False, stats=False):
        org = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, dtype=np.float32).reshape(-1, 3)
        n = org.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(tmax, dtype=np.float32)
        prim = np.empty(n, np.int32); inst = np.empty(n, np.int32)
        t = np.empty(n, np.float32); u = np.empty(n, np.float32); v = np.empty(n, np.float32)
        if brute:
            self._L.orc_trace_brute(self._h, _fp(org), _fp(dir), _fp(tm), n, int(any_hit),
                                    prim.ctypes.data_as(_i32), inst.ctypes.data_as(_i32), _fp(t), _fp(u), _fp(v))
            return prim, inst, t, u, v
        st = crt_stats()
        self._L.orc_trace(self._h, _fp(org), _fp(dir), _fp(tm), n, int(any_hit),
                          prim.ctypes.data_as(_i32), inst.ctypes.data_as(_i32), _fp(t), _fp(u), _fp(v),
                          C.byref(st) if stats else None)
        if stats:
            return prim, inst, t, u, v, st.as_dict()
        return prim, inst, t, u, v

    def render(self, w, h, n_samples, first_sample=0, accum=None, nthreads=0, stats=False):
        if accum is None:
            accum = np.zeros((h, w, 4), dtype=np.float32)
        st = crt_stats()
        self._L.orc_render(self._h, w, h, first_sample, n_samples, _fp(accum), nthreads, C.byref(st) if stats else None)
        return (accum, st.as_dict()) if stats else accum

    def render_adaptive(self, w, h, tile_samples, wave_cap, state=None, first_sample=0, nthreads=0):
        """Adaptive screen sampling (orc_render_adaptive).  `state` carries accum / tile counts / tile errors /
        even-sample luminance / wave index between calls; returns it."""
        nt = ((w + 31) // 32) * ((h + 31) // 32)
        if state is None:
            state = {"accum": np.zeros((h, w, 4), np.float32), "count": np.zeros(nt, np.uint32),
                     "err": np.zeros(nt, np.uint32), "even": np.zeros((h, w), np.float32), "wave": C.c_uint32(0)}
        _u32 = C.POINTER(C.c_uint32)
        self._L.orc_render_adaptive(self._h, w, h, first_sample, tile_samples, wave_cap, _fp(state["accum"]),
                                    state["count"].ctypes.data_as(_u32), state["err"].ctypes.data_as(_u32),
                                    _fp(state["even"]), C.byref(state["wave"]), nthreads)
        return state

    def display(self, accum):
        h, w = accum.shape[:2]
        out = np.empty((h, w, 3), dtype=np.uint8)
        self._L.orc_display(self._h, _fp(accum), w, h, out.ctypes.data_as(_u8))
        return out

    def hdr(self, accum):
        h, w = accum.shape[:2]
        out = np.empty((h, w, 3), dtype=np.float32)
        self._L.orc_hdr(_fp(accum), w, h, _fp(out))
        return out

    def epsilon(self) -> float:
        return float(self._L.orc_scene_epsilon(self._h))


def adaptive_allocate(tile_err, budget, wave):
    """orc_adaptive_allocate: exclusive prefix (nt + 1 entries) of the tile samples a wave hands to each tile."""
    err = np.ascontiguousarray(tile_err, dtype=np.uint32)
    cum = np.zeros(err.size + 1, np.uint32)
    _u32 = C.POINTER(C.c_uint32)
    lib().orc_adaptive_allocate(err.ctypes.data_as(_u32), err.size, int(budget), int(wave), cum.ctypes.data_as(_u32))
    return cum


if __name__ == "__main__":
    print("built", build_oracle(force="--force" in sys.argv))
