"""Host-side mirror of the OCCT interface CADRays drives for this path.

Same names, argument meaning and error behaviour as the reference call sites
(file:line relative to the CADRays repository):

  Graphic3d_Fresnel / Graphic3d_BSDF   src/Launcher/MaterialEditor.cxx:177-201,281-331
  Graphic3d_RenderingParams            src/Launcher/SettingsWidget.cxx:65-90,217-229,263-478
  V3d_View::Redraw / BufferDump        src/Launcher/AppViewer.cxx:1047,1259-1262; AppGui.cxx:345-350,430
  V3d lights / SetTextureEnv           src/Launcher/LightSourcesEditor.cxx:242-369

Everything computes in libcadrays_b200.so (sm_100a kernels); this file only
marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import check, crt_bsdf, crt_camera, crt_light, crt_params, crt_stats

# Graphic3d_RenderingMethod / Graphic3d_ToneMappingMethod / Graphic3d_BufferType
Graphic3d_RM_RASTERIZATION = 0
Graphic3d_RM_RAYTRACING = 1
Graphic3d_ToneMappingMethod_Disabled = 0
Graphic3d_ToneMappingMethod_Filmic = 1
Graphic3d_BT_RGB = 0
Graphic3d_BT_RGB_RayTraceHdrLeft = 1

Graphic3d_FM_SCHLICK, Graphic3d_FM_CONSTANT, Graphic3d_FM_CONDUCTOR, Graphic3d_FM_DIELECTRIC = 0, 1, 2, 3


def _clamp(x, lo, hi):
    return min(max(float(x), lo), hi)


class Graphic3d_Fresnel:
    """Interface model of one BSDF layer; Serialize() is the vec4 the shader reads
    (MaterialEditor.cxx:209-255, ImportExport.cxx:204-227)."""

    def __init__(self, ftype: int, data: Sequence[float]):
        self._type = ftype
        self._data = tuple(float(v) for v in data)

    @staticmethod
    def CreateSchlick(r, g=None, b=None):
        if g is None:
            r, g, b = r
        return Graphic3d_Fresnel(Graphic3d_FM_SCHLICK, (_clamp(r, 0, 1), _clamp(g, 0, 1), _clamp(b, 0, 1)))

    @staticmethod
    def CreateConstant(reflection):
        return Graphic3d_Fresnel(Graphic3d_FM_CONSTANT, (0.0, _clamp(reflection, 0, 1), 0.0))

    @staticmethod
    def CreateConductor(n, k):
        return Graphic3d_Fresnel(Graphic3d_FM_CONDUCTOR, (float(n), float(k), 0.0))

    @staticmethod
    def CreateDielectric(ior):
        return Graphic3d_Fresnel(Graphic3d_FM_DIELECTRIC, (float(ior), 0.0, 0.0))

    def FresnelType(self) -> int:
        return self._type

    def Serialize(self):
        d = self._data
        if self._type == Graphic3d_FM_SCHLICK:
            return (d[0], d[1], d[2], 0.0)
        if self._type == Graphic3d_FM_CONSTANT:
            return (-1.0, 0.0, d[1], 0.0)
        if self._type == Graphic3d_FM_CONDUCTOR:
            return (-2.0, d[0], d[1], 0.0)
        return (-3.0, d[0], 0.0, 0.0)


@dataclass
class Graphic3d_BSDF:
    """The fields CADRays edits (MaterialEditor.cxx:294-331)."""
    Kc: list = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])   # rgb + coat roughness
    Kd: list = field(default_factory=lambda: [0.0, 0.0, 0.0])
    Ks: list = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])   # rgb + base roughness
    Kt: list = field(default_factory=lambda: [0.0, 0.0, 0.0])
    Le: list = field(default_factory=lambda: [0.0, 0.0, 0.0])
    Absorption: list = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])  # rgb colour + coefficient
    FresnelCoat: Graphic3d_Fresnel = field(default_factory=lambda: Graphic3d_Fresnel.CreateConstant(0.0))
    FresnelBase: Graphic3d_Fresnel = field(default_factory=lambda: Graphic3d_Fresnel.CreateConstant(1.0))
    # texture map of the aspect (SetTextureMap / SetTextureMapOn, AisMesh.cxx:343-345): index into the
    # view's texture list or None, and the "rttexture -scale S T" factors
    TextureId: Optional[int] = None
    TextureScale: tuple = (1.0, 1.0)

    @staticmethod
    def CreateDiffuse(weight):
        return Graphic3d_BSDF(Kd=list(weight))

    @staticmethod
    def CreateMetallic(weight, fresnel: Graphic3d_Fresnel, roughness: float):
        return Graphic3d_BSDF(Ks=[*weight, float(roughness)], FresnelBase=fresnel)

    @staticmethod
    def CreateTransparent(weight, absorption_color, absorption_coeff):
        return Graphic3d_BSDF(Kt=list(weight), Absorption=[*absorption_color, float(absorption_coeff)])

    @staticmethod
    def CreateGlass(weight, absorption_color, absorption_coeff, ior):
        b = Graphic3d_BSDF(Kt=list(weight), Absorption=[*absorption_color, float(absorption_coeff)])
        b.Kc = [1.0, 1.0, 1.0, 0.0]
        b.FresnelCoat = Graphic3d_Fresnel.CreateDielectric(ior)
        return b

    def Normalize(self):
        """CADRays' clamp + base-layer normalisation (MaterialEditor.cxx:294-329)."""
        for v in (self.Kc, self.Kd, self.Ks, self.Kt, self.Absorption):
            for k in range(3):
                v[k] = _clamp(v[k], 0.0, 1.0)
        for k in range(3):
            self.Le[k] = max(float(self.Le[k]), 0.0)
        self.Absorption[3] = max(float(self.Absorption[3]), 0.0)
        mx = max(self.Kd[k] + self.Ks[k] + self.Kt[k] for k in range(3))
        if mx > 1.0:
            for k in range(3):
                self.Kd[k] /= mx
                self.Ks[k] /= mx
                self.Kt[k] /= mx
        return self

    def to_c(self) -> crt_bsdf:
        c = crt_bsdf()
        c.Kc[:] = self.Kc
        tex = 0.0 if self.TextureId is None else float(self.TextureId + 1)
        c.Kd[:] = [*self.Kd, tex]
        c.Ks[:] = self.Ks
        c.Kt[:] = [*self.Kt, float(self.TextureScale[0]) if tex else 0.0]
        c.Le[:] = [*self.Le, float(self.TextureScale[1]) if tex else 0.0]
        c.FresnelCoat[:] = self.FresnelCoat.Serialize()
        c.FresnelBase[:] = self.FresnelBase.Serialize()
        c.Absorption[:] = self.Absorption
        return c


@dataclass
class Graphic3d_RenderingParams:
    """Fields of Graphic3d_RenderingParams that CADRays sets."""
    Method: int = Graphic3d_RM_RAYTRACING
    IsGlobalIlluminationEnabled: bool = True
    RaytracingDepth: int = 8
    SamplesPerPixel: int = 1
    RadianceClampingValue: float = 50.0
    TwoSidedBsdfModels: bool = False
    CoherentPathTracingMode: bool = False
    AdaptiveScreenSampling: bool = False        # SettingsWidget.cxx:70,427-436; vrenderparams -iss
    NbRayTracingTiles: int = 128                # CADRays' default (SettingsWidget.cxx:72); the GUI writes 64..1024 (:471-476)
    ShowSamplingTiles: bool = False             # debug view (SettingsWidget.cxx:443-449): V3d_View.SamplingTiles()
    ToneMappingMethod: int = Graphic3d_ToneMappingMethod_Disabled
    WhitePoint: float = 1.0
    Exposure: float = 0.0
    CameraApertureRadius: float = 0.0
    CameraFocalPlaneDist: float = 1.0
    UseEnvironmentMapBackground: bool = True
    # not OCCT fields: generator seed, roulette switch, background colour, wave size
    FrameSeed: int = 1
    RussianRoulette: bool = True
    BackgroundColor: tuple = (0.0, 0.0, 0.0)
    SamplesPerBatch: int = 0
    BvhWidth: int = 2             # 4 = OCCT's optional QUAD_BVH collapse (scene is rebuilt at the next Update)

    def to_c(self) -> crt_params:
        if self.Method != Graphic3d_RM_RAYTRACING or not self.IsGlobalIlluminationEnabled:
            raise ValueError("only Graphic3d_RM_RAYTRACING with IsGlobalIlluminationEnabled is implemented "
                             "(the path-traced mode CADRays calls 'GI', SettingsWidget.cxx:76-84)")
        p = crt_params()
        p.max_depth = int(self.RaytracingDepth)
        p.max_radiance = float(self.RadianceClampingValue)
        p.two_sided = int(self.TwoSidedBsdfModels)
        p.coherent_rng = int(self.CoherentPathTracingMode)
        p.aperture_radius = float(self.CameraApertureRadius)
        p.focal_dist = float(self.CameraFocalPlaneDist)
        p.tone_map = int(self.ToneMappingMethod)
        p.white_point = float(self.WhitePoint)
        p.exposure = float(self.Exposure)
        p.env_as_background = int(self.UseEnvironmentMapBackground)
        p.frame_seed0 = int(self.FrameSeed) & 0xFFFFFFFF
        p.russian_roulette = int(self.RussianRoulette)
        p.background[:] = [float(v) for v in self.BackgroundColor]
        p.samples_per_batch = int(self.SamplesPerBatch)
        p.bvh_width = int(self.BvhWidth)
        p.adaptive_sampling = int(self.AdaptiveScreenSampling)
        p.adaptive_tiles = int(self.NbRayTracingTiles)
        return p


@dataclass
class Graphic3d_Camera:
    Eye: tuple = (0.0, -3.0, 0.0)
    Direction: tuple = (0.0, 1.0, 0.0)
    Up: tuple = (0.0, 0.0, 1.0)
    FOVy: float = 45.0
    Aspect: float = 1.0
    IsOrthographic: bool = False
    Scale: float = 1.0

    def to_c(self) -> crt_camera:
        c = crt_camera()
        c.eye[:] = [float(v) for v in self.Eye]
        c.dir[:] = [float(v) for v in self.Direction]
        c.up[:] = [float(v) for v in self.Up]
        c.fovy_deg = float(self.FOVy)
        c.aspect = float(self.Aspect)
        c.is_ortho = int(self.IsOrthographic)
        c.ortho_scale = float(self.Scale)
        return c


def make_light(is_point: bool, posdir, color=(1.0, 1.0, 1.0), intensity=1.0, smoothness=0.0) -> crt_light:
    """V3d_PositionalLight / V3d_DirectionalLight (LightSourcesEditor.cxx:242-310)."""
    l = crt_light()
    l.emission[:] = [float(c) * float(intensity) for c in color]
    l.smoothness = float(smoothness)
    l.posdir[:] = [float(v) for v in posdir]
    l.is_point = int(bool(is_point))
    return l


def _fptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class V3d_View:
    """One render target: owns a crt_context on `device`.  With `devices=[d0, d1, ...]` (d0 = the view's own
    device) the same view drives every listed GPU from this one process through a crt_group: Update(), Redraw(),
    BufferDump() and ResetAccumulation() go to the group, everything else (scene, materials, lights, camera,
    parameters) is set on the view as before and replicated by the library."""

    def __init__(self, device: int = 0, host_only: bool = False, devices: Optional[Sequence[int]] = None):
        self._lib = _ffi.load_library()
        self._ctx = C.c_void_p()
        self._group = C.c_void_p()
        if devices is not None and len(devices):
            device = int(devices[0])
        if host_only:
            check(self._lib.crt_create_host_only(C.byref(self._ctx)))
        else:
            check(self._lib.crt_create(int(device), C.byref(self._ctx)))
        if devices is not None and len(devices) and not host_only:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            try:
                check(self._lib.crt_group_create(self._ctx, arr, len(devices), C.byref(self._group)))
            except Exception:
                self._lib.crt_destroy(self._ctx)
                self._ctx = C.c_void_p()
                raise
        self._params = Graphic3d_RenderingParams()
        self._camera = Graphic3d_Camera()
        self._size = (0, 0)
        self.device = device
        self.devices = [int(d) for d in devices] if devices is not None and len(devices) else [int(device)]

    # -- lifetime
    def Remove(self):
        if self._group:
            self._lib.crt_group_destroy(self._group)
            self._group = C.c_void_p()
        if self._ctx:
            self._lib.crt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def GroupInfo(self) -> dict:
        """crt_group_info: peer access, NCCL path in use, device ms of the last exchange + Display pass."""
        if not self._group:
            return {"members": 1, "peer_access": True, "nccl": False, "reduce_ms": 0.0}
        a, b, ms = C.c_int(), C.c_int(), C.c_double()
        check(self._lib.crt_group_info(self._group, C.byref(a), C.byref(b), C.byref(ms)))
        return {"members": self._lib.crt_group_size(self._group), "peer_access": bool(a.value), "nccl": bool(b.value),
                "reduce_ms": ms.value}

    def MemberHandle(self, rank: int):
        if not self._group:
            return self._ctx
        out = C.c_void_p()
        check(self._lib.crt_group_member(self._group, int(rank), C.byref(out)))
        return out

    def CommitStats(self) -> int:
        """How many Update() calls re-uploaded the top-level nodes + instance records only."""
        n = C.c_uint64()
        check(self._lib.crt_commit_stats(self._ctx, C.byref(n)))
        return n.value

    def __del__(self):
        try:
            self.Remove()
        except Exception:
            pass

    @property
    def handle(self):
        return self._ctx

    # -- scene (AIS Display of a triangulation with a location and a material aspect)
    def AddMesh(self, pos, idx, nrm=None, uv=None) -> int:
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        nrm = None if nrm is None else np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        # crt_mesh_create reads 3 * n_verts normals and 2 * n_verts texels: refuse attribute arrays of another length
        if nrm is not None and nrm.shape[0] != pos.shape[0]:
            raise ValueError(f"normals: {nrm.shape[0]} rows for {pos.shape[0]} vertices")
        if uv is not None and uv.shape[0] != pos.shape[0]:
            raise ValueError(f"texels: {uv.shape[0]} rows for {pos.shape[0]} vertices")
        out = C.c_uint32()
        check(self._lib.crt_mesh_create(self._ctx, _fptr(pos), _fptr(nrm), _fptr(uv), pos.shape[0],
                                        idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.shape[0], C.byref(out)))
        return out.value

    def Display(self, mesh_id: int, trsf=None, material_id: int = 0) -> int:
        xf = None if trsf is None else np.ascontiguousarray(trsf, dtype=np.float32).reshape(12)
        out = C.c_uint32()
        check(self._lib.crt_instance_add(self._ctx, int(mesh_id), _fptr(xf), int(material_id), C.byref(out)))
        return out.value

    def SetLocation(self, inst_id: int, trsf):
        xf = None if trsf is None else np.ascontiguousarray(trsf, dtype=np.float32).reshape(12)   # None = identity
        check(self._lib.crt_instance_set_transform(self._ctx, int(inst_id), _fptr(xf)))

    def SetMaterialIndex(self, inst_id: int, material_id: int):
        check(self._lib.crt_instance_set_material(self._ctx, int(inst_id), int(material_id)))

    def SetVisible(self, inst_id: int, visible: bool):
        """AIS Erase / Display of an object that stays in the scene (takes effect at the next Update())."""
        check(self._lib.crt_instance_set_visible(self._ctx, int(inst_id), int(bool(visible))))

    def Clear(self):
        check(self._lib.crt_scene_clear(self._ctx))

    def SetMaterials(self, bsdfs: Sequence):
        arr = (crt_bsdf * max(len(bsdfs), 1))()
        for i, b in enumerate(bsdfs):
            arr[i] = b.to_c() if isinstance(b, Graphic3d_BSDF) else b
        check(self._lib.crt_materials_set(self._ctx, arr, len(bsdfs)))

    def AddTexture(self, image: np.ndarray) -> int:
        """Graphic3d_Texture2Dmanual: uint8 (h, w, 3|4), rows top-down as in the image file."""
        img = np.ascontiguousarray(image, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] not in (3, 4):
            raise ValueError("texture must be (h, w, 3) or (h, w, 4) uint8")
        if img.shape[2] == 3:
            img = np.ascontiguousarray(np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2))
        out = C.c_uint32()
        check(self._lib.crt_texture_create(self._ctx, img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[1], img.shape[0], C.byref(out)))
        return out.value

    def ClearTextures(self):
        check(self._lib.crt_textures_clear(self._ctx))

    def SetLights(self, lights: Sequence[crt_light]):
        arr = (crt_light * max(len(lights), 1))()
        for i, l in enumerate(lights):
            arr[i] = l
        check(self._lib.crt_lights_set(self._ctx, arr, len(lights)))

    def SetTextureEnv(self, image: Optional[np.ndarray]):
        """Lat-long environment map: uint8 (h,w,3) is linearised as (c/255)^2, float32 is linear."""
        if image is None:
            check(self._lib.crt_envmap_set_rgb32f(self._ctx, None, 0, 0))
            return
        h, w = image.shape[:2]
        if image.dtype == np.uint8:
            a = np.ascontiguousarray(image[..., :3])
            check(self._lib.crt_envmap_set_rgb8(self._ctx, a.ctypes.data_as(C.POINTER(C.c_uint8)), w, h))
        else:
            a = np.ascontiguousarray(image[..., :3], dtype=np.float32)
            check(self._lib.crt_envmap_set_rgb32f(self._ctx, _fptr(a), w, h))

    # -- parameters
    def ChangeRenderingParams(self) -> Graphic3d_RenderingParams:
        return self._params

    def SetRenderingParams(self, p: Graphic3d_RenderingParams):
        self._params = p
        cp = p.to_c()
        check(self._lib.crt_params_set(self._ctx, C.byref(cp)))

    def Camera(self) -> Graphic3d_Camera:
        return self._camera

    def SetCamera(self, cam: Graphic3d_Camera):
        self._camera = cam
        cc = cam.to_c()
        check(self._lib.crt_camera_set(self._ctx, C.byref(cc)))

    def SetWindowSize(self, width: int, height: int):
        check(self._lib.crt_resize(self._ctx, int(width), int(height)))
        self._size = (int(width), int(height))

    def Update(self):
        """Explicit form of OCCT's state-counter invalidation: BVH (re)build + upload."""
        if self._group:
            check(self._lib.crt_group_commit(self._group))
        else:
            check(self._lib.crt_commit(self._ctx))

    # -- render
    def Redraw(self, samples: Optional[int] = None) -> int:
        """V3d_View::Redraw(): adds SamplesPerPixel samples (or `samples`); returns the total."""
        n = int(self._params.SamplesPerPixel if samples is None else samples)
        total = C.c_uint64()
        if self._group:
            check(self._lib.crt_group_render(self._group, n, C.byref(total)))
        else:
            check(self._lib.crt_render(self._ctx, n, C.byref(total)))
        return total.value

    def RedrawAsync(self, samples: int):
        check(self._lib.crt_render_async(self._ctx, int(samples)))

    def Sync(self):
        check(self._lib.crt_sync(self._ctx))

    def ResetAccumulation(self, first_sample: int = 0):
        if self._group:
            check(self._lib.crt_group_reset_accumulation(self._group, int(first_sample)))
        else:
            check(self._lib.crt_reset_accumulation(self._ctx, int(first_sample)))

    def SetNextSample(self, index: int):
        check(self._lib.crt_set_next_sample(self._ctx, int(index)))

    def BufferDump(self, buffer_type: int = Graphic3d_BT_RGB, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Graphic3d_CView::BufferDump: RGB8 (tone-mapped) or float RGB (mean radiance), bottom-up rows."""
        w, h = self._size
        dtype = np.uint8 if buffer_type == Graphic3d_BT_RGB else np.float32
        if out is not None:
            # the library writes h * w * 3 elements through this pointer
            if not isinstance(out, np.ndarray) or out.dtype != dtype or out.shape != (h, w, 3) or not out.flags["C_CONTIGUOUS"] \
                    or not out.flags["WRITEABLE"]:
                raise ValueError(f"out must be a writable C-contiguous {np.dtype(dtype).name} array of shape {(h, w, 3)}")
        img = out if out is not None else np.empty((h, w, 3), dtype=dtype)
        if self._group:
            if buffer_type == Graphic3d_BT_RGB:
                check(self._lib.crt_group_read_ldr(self._group, img.ctypes.data_as(C.POINTER(C.c_uint8)), 0))
            else:
                check(self._lib.crt_group_read_hdr(self._group, _fptr(img), 0))
        elif buffer_type == Graphic3d_BT_RGB:
            check(self._lib.crt_read_ldr(self._ctx, img.ctypes.data_as(C.POINTER(C.c_uint8)), 0))
        else:
            check(self._lib.crt_read_hdr(self._ctx, _fptr(img), 0))
        return img

    def SamplingTiles(self):
        """Adaptive screen sampling state (what Graphic3d_RenderingParams::ShowSamplingTiles displays):
        (samples per pixel, error estimate) of every 32x32 tile, each of shape (tiles_y, tiles_x)."""
        tx, ty = C.c_uint32(), C.c_uint32()
        check(self._lib.crt_adaptive_tiles_get(self._ctx, None, None, 0, C.byref(tx), C.byref(ty)))
        n = tx.value * ty.value
        counts, errs = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        u32 = C.POINTER(C.c_uint32)
        check(self._lib.crt_adaptive_tiles_get(self._ctx, counts.ctypes.data_as(u32), errs.ctypes.data_as(u32), n, None, None))
        return counts.reshape(ty.value, tx.value), errs.reshape(ty.value, tx.value)

    def ToPixMap(self, width: int, height: int, buffer_type: int = Graphic3d_BT_RGB, samples: Optional[int] = None) -> np.ndarray:
        """V3d_View::ToPixMap(Image_PixMap&, width, height, bufferType): an off-screen render at the given size --
        resize, render `samples` (default SamplesPerPixel) samples per pixel, dump.  CADRays itself uses BufferDump on
        the live view (AppViewer.cxx:1259-1262); ToPixMap is what DRAW's `vdump -width -height` goes through."""
        import copy
        old_size, old_cam = self._size, copy.copy(self._camera)
        self.SetWindowSize(int(width), int(height))
        cam = copy.copy(self._camera)
        cam.Aspect = float(width) / float(height)
        self.SetCamera(cam)
        self.Redraw(samples)
        img = self.BufferDump(buffer_type)
        # OCCT's ToPixMap renders into its own FBO and leaves the view as it was: restore window size and camera
        # (the live accumulation restarts, as it does after any resize)
        if old_size[0] and old_size[1] and old_size != (int(width), int(height)):
            self.SetWindowSize(*old_size)
        self.SetCamera(old_cam)
        return img

    def AccumDevicePtr(self):
        p = C.c_void_p()
        n = C.c_size_t()
        check(self._lib.crt_accum_device_ptr(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def BindAccum(self, device_ptr: Optional[int], nbytes: int = 0):
        check(self._lib.crt_accum_bind(self._ctx, C.c_void_p(device_ptr) if device_ptr else None, nbytes))

    def DumpFrom(self, device_ptr: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        w, h = self._size
        if out is not None and (not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != (h, w, 3)
                                or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]):
            raise ValueError(f"out must be a writable C-contiguous uint8 array of shape {(h, w, 3)}")
        img = out if out is not None else np.empty((h, w, 3), dtype=np.uint8)
        check(self._lib.crt_read_ldr_from(self._ctx, C.c_void_p(device_ptr), img.ctypes.data_as(C.POINTER(C.c_uint8)), 0))
        return img

    # -- parity hooks
    def Trace(self, org, dir, tmax=None, any_hit: bool = False):
        org = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, dtype=np.float32).reshape(-1, 3)
        n = org.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(tmax, dtype=np.float32).reshape(n)
        prim = np.empty(n, np.int32); inst = np.empty(n, np.int32)
        t = np.empty(n, np.float32); u = np.empty(n, np.float32); v = np.empty(n, np.float32)
        ip = C.POINTER(C.c_int32)
        check(self._lib.crt_trace(self._ctx, _fptr(org), _fptr(dir), _fptr(tm), n, int(any_hit),
                                  prim.ctypes.data_as(ip), inst.ctypes.data_as(ip), _fptr(t), _fptr(u), _fptr(v)))
        return prim, inst, t, u, v

    def WavefrontRays(self, depth: int, shadow: bool = False):
        """crt_wavefront_rays: (org, dir, tmax) of the rays the last wave traced at bounce `depth` -- valid when the
        wave ran with RaytracingDepth == depth + 1 (see the header)."""
        n = C.c_uint32()
        check(self._lib.crt_wavefront_rays(self._ctx, int(depth), int(bool(shadow)), None, None, None, 0, C.byref(n)))
        org = np.empty((n.value, 3), np.float32); d = np.empty((n.value, 3), np.float32); tm = np.empty(n.value, np.float32)
        if n.value:
            check(self._lib.crt_wavefront_rays(self._ctx, int(depth), int(bool(shadow)), _fptr(org), _fptr(d), _fptr(tm),
                                               n.value, C.byref(n)))
        return org, d, tm

    def TraceDevice(self, org4_ptr: int, dir4_ptr: int, n: int, hit4_ptr: int, inst_ptr: int = 0, any_hit: bool = False):
        check(self._lib.crt_trace_device(self._ctx, C.c_void_p(org4_ptr), C.c_void_p(dir4_ptr), int(n), int(any_hit),
                                         C.c_void_p(hit4_ptr), C.c_void_p(inst_ptr) if inst_ptr else None))

    def ExportBVH(self) -> bytes:
        n = C.c_size_t()
        check(self._lib.crt_bvh_export(self._ctx, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        check(self._lib.crt_bvh_export(self._ctx, buf, n.value, C.byref(n)))
        return buf.raw

    def ImportBVH(self, blob: bytes):
        check(self._lib.crt_bvh_import(self._ctx, blob, len(blob)))

    # -- metrics
    def EnableStats(self, on: bool = True):
        check(self._lib.crt_stats_enable(self._ctx, int(on)))

    def ResetStats(self):
        check(self._lib.crt_stats_reset(self._ctx))

    def Stats(self) -> dict:
        s = crt_stats()
        check(self._lib.crt_stats_get(self._ctx, C.byref(s)))
        return s.as_dict()

    def EnableTiming(self, on: bool = True):
        check(self._lib.crt_timing_enable(self._ctx, int(on)))

    def Timing(self):
        ms = (C.c_double * 6)()
        ln = (C.c_uint64 * 6)()
        check(self._lib.crt_timing_get(self._ctx, ms, ln))
        names = ("generate", "extend", "shade", "connect", "resolve", "render")
        return {n: (ms[i], int(ln[i])) for i, n in enumerate(names)}

    def LaunchCount(self) -> int:
        """Kernels enqueued by this view's context since ResetStats()."""
        n = C.c_uint64()
        check(self._lib.crt_launch_count(self._ctx, C.byref(n)))
        return int(n.value)

    def SceneBytes(self):
        """(bytes the traversal kernels read, bytes of the whole committed scene) in device memory."""
        a, b = C.c_size_t(), C.c_size_t()
        check(self._lib.crt_scene_bytes(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def Stream(self) -> int:
        p = C.c_void_p()
        check(self._lib.crt_stream(self._ctx, C.byref(p)))
        return p.value or 0
