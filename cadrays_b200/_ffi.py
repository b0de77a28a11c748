"""ctypes binding of libcadrays_b200.so (include/cadrays_b200.h).

The library is the product; this module only declares its C-ABI.  There is no
fallback of any kind: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libcadrays_b200.so"

CRT_OK = 0
CRT_ERR_INVALID_ARG = -1
CRT_ERR_NO_DEVICE = -2
CRT_ERR_CUDA = -3
CRT_ERR_OUT_OF_MEMORY = -4
CRT_ERR_STATE = -5
CRT_ERR_FORMAT = -6


class CrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libcadrays_b200 error {code}: {message}")
        self.code = code
        self.message = message


class crt_bsdf(C.Structure):
    _fields_ = [(n, C.c_float * 4) for n in
                ("Kc", "Kd", "Ks", "Kt", "Le", "FresnelCoat", "FresnelBase", "Absorption")]


class crt_light(C.Structure):
    _fields_ = [("emission", C.c_float * 3), ("smoothness", C.c_float),
                ("posdir", C.c_float * 3), ("is_point", C.c_int32)]


class crt_params(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("max_radiance", C.c_float), ("two_sided", C.c_int32),
                ("coherent_rng", C.c_int32), ("aperture_radius", C.c_float), ("focal_dist", C.c_float),
                ("tone_map", C.c_int32), ("white_point", C.c_float), ("exposure", C.c_float),
                ("env_as_background", C.c_int32), ("frame_seed0", C.c_uint32),
                ("russian_roulette", C.c_int32), ("background", C.c_float * 3),
                ("samples_per_batch", C.c_int32), ("bvh_width", C.c_int32),
                ("adaptive_sampling", C.c_int32), ("adaptive_tiles", C.c_int32)]


class crt_camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("dir", C.c_float * 3), ("up", C.c_float * 3),
                ("fovy_deg", C.c_float), ("aspect", C.c_float), ("is_ortho", C.c_int32),
                ("ortho_scale", C.c_float)]


class crt_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("rays_nearest", "rays_any", "n_inner", "n_leaf", "n_tri", "n_switch", "shaded_hits", "samples",
                 "n_inner_any", "n_leaf_any", "n_tri_any", "n_switch_any", "n_boxes", "n_boxes_any")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_ctx = C.c_void_p
_f = C.POINTER(C.c_float)
_u32 = C.POINTER(C.c_uint32)
_i32 = C.POINTER(C.c_int32)
_u8 = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol include/cadrays_b200.h declares
PROTOTYPES = {
    "crt_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "crt_create_host_only": (C.c_int, [C.POINTER(_ctx)]),
    "crt_destroy": (None, [_ctx]),
    "crt_last_error": (C.c_char_p, []),
    "crt_abi_version": (C.c_int, []),
    "crt_mesh_create": (C.c_int, [_ctx, _f, _f, _f, C.c_uint32, _u32, C.c_uint32, _u32]),
    "crt_instance_add": (C.c_int, [_ctx, C.c_uint32, _f, C.c_uint32, _u32]),
    "crt_instance_set_transform": (C.c_int, [_ctx, C.c_uint32, _f]),
    "crt_instance_set_material": (C.c_int, [_ctx, C.c_uint32, C.c_uint32]),
    "crt_instance_set_visible": (C.c_int, [_ctx, C.c_uint32, C.c_int]),
    "crt_scene_clear": (C.c_int, [_ctx]),
    "crt_materials_set": (C.c_int, [_ctx, C.POINTER(crt_bsdf), C.c_uint32]),
    "crt_texture_create": (C.c_int, [_ctx, _u8, C.c_uint32, C.c_uint32, _u32]),
    "crt_textures_clear": (C.c_int, [_ctx]),
    "crt_lights_set": (C.c_int, [_ctx, C.POINTER(crt_light), C.c_uint32]),
    "crt_envmap_set_rgb8": (C.c_int, [_ctx, _u8, C.c_uint32, C.c_uint32]),
    "crt_envmap_set_rgb32f": (C.c_int, [_ctx, _f, C.c_uint32, C.c_uint32]),
    "crt_params_default": (C.c_int, [C.POINTER(crt_params)]),
    "crt_params_set": (C.c_int, [_ctx, C.POINTER(crt_params)]),
    "crt_camera_set": (C.c_int, [_ctx, C.POINTER(crt_camera)]),
    "crt_resize": (C.c_int, [_ctx, C.c_uint32, C.c_uint32]),
    "crt_commit": (C.c_int, [_ctx]),
    "crt_render": (C.c_int, [_ctx, C.c_uint32, C.POINTER(C.c_uint64)]),
    "crt_adaptive_tiles_get": (C.c_int, [_ctx, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32,
                                         C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "crt_render_async": (C.c_int, [_ctx, C.c_uint32]),
    "crt_sync": (C.c_int, [_ctx]),
    "crt_reset_accumulation": (C.c_int, [_ctx, C.c_uint64]),
    "crt_set_next_sample": (C.c_int, [_ctx, C.c_uint64]),
    "crt_read_ldr": (C.c_int, [_ctx, _u8, C.c_size_t]),
    "crt_read_hdr": (C.c_int, [_ctx, _f, C.c_size_t]),
    "crt_read_ldr_from": (C.c_int, [_ctx, C.c_void_p, _u8, C.c_size_t]),
    "crt_accum_device_ptr": (C.c_int, [_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "crt_accum_bind": (C.c_int, [_ctx, C.c_void_p, C.c_size_t]),
    "crt_trace": (C.c_int, [_ctx, _f, _f, _f, C.c_uint32, C.c_int, _i32, _i32, _f, _f, _f]),
    "crt_wavefront_rays": (C.c_int, [_ctx, C.c_int, C.c_int, _f, _f, _f, C.c_uint32, C.POINTER(C.c_uint32)]),
    "crt_trace_device": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "crt_bvh_export": (C.c_int, [_ctx, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "crt_bvh_import": (C.c_int, [_ctx, C.c_void_p, C.c_size_t]),
    "crt_stats_enable": (C.c_int, [_ctx, C.c_int]),
    "crt_stats_reset": (C.c_int, [_ctx]),
    "crt_stats_get": (C.c_int, [_ctx, C.POINTER(crt_stats)]),
    "crt_timing_enable": (C.c_int, [_ctx, C.c_int]),
    "crt_timing_get": (C.c_int, [_ctx, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "crt_launch_count": (C.c_int, [_ctx, C.POINTER(C.c_uint64)]),
    "crt_scene_bytes": (C.c_int, [_ctx, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "crt_stream": (C.c_int, [_ctx, C.POINTER(C.c_void_p)]),
    "crt_commit_stats": (C.c_int, [_ctx, C.POINTER(C.c_uint64)]),
    "crt_group_create": (C.c_int, [_ctx, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "crt_group_destroy": (None, [C.c_void_p]),
    "crt_group_size": (C.c_int, [C.c_void_p]),
    "crt_group_member": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_ctx)]),
    "crt_group_commit": (C.c_int, [C.c_void_p]),
    "crt_group_render": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "crt_group_reset_accumulation": (C.c_int, [C.c_void_p, C.c_uint64]),
    "crt_group_read_ldr": (C.c_int, [C.c_void_p, _u8, C.c_size_t]),
    "crt_group_read_hdr": (C.c_int, [C.c_void_p, _f, C.c_size_t]),
    "crt_group_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
}

_lib = None


def load_library() -> C.CDLL:
    """Loads the in-tree library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("CADRAYS_B200_LIB", LIB_PATH))     # A/B variants of the same library
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m cadrays_b200.build` "
            "(there is no CPU or PyTorch fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != CRT_OK:
        msg = load_library().crt_last_error()
        raise CrtError(code, msg.decode("utf-8", "replace") if msg else "")
