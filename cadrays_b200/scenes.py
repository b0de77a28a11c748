"""Procedural scenes for the BASELINE.json configs.

Own tessellators (box / quad / UV sphere / cylinder / torus / L-bracket) and a
PCG32 generator with fixed seeds; nothing here computes rendering.  C1 and C3
mirror data/scripts/CornellBox.tcl and data/scripts/Materials.tcl of the
reference (line numbers cited at each object).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .view import (Graphic3d_BSDF, Graphic3d_Camera, Graphic3d_Fresnel, Graphic3d_RenderingParams,
                   V3d_View, make_light)


class PCG32:
    """PCG-XSH-RR 64/32 (O'Neill); scene generation only."""

    def __init__(self, seed: int, seq: int = 54):
        self.state = 0
        self.inc = ((seq << 1) | 1) & 0xFFFFFFFFFFFFFFFF
        self.next_u32()
        self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF
        self.next_u32()

    def next_u32(self) -> int:
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * (self.next_u32() / 4294967296.0)

    def randint(self, n: int) -> int:
        return self.next_u32() % n


# ------------------------------------------------------------------ tessellators
# each returns (pos float32 (n,3), nrm float32 (n,3), idx uint32 (m,3)); counter-clockwise seen from outside

def _grid_face(origin, eu, ev, normal, n):
    u = np.linspace(0.0, 1.0, n + 1, dtype=np.float64)
    uu, vv = np.meshgrid(u, u, indexing="xy")
    p = origin[None, None, :] + uu[..., None] * eu[None, None, :] + vv[..., None] * ev[None, None, :]
    pos = p.reshape(-1, 3)
    nrm = np.broadcast_to(normal, pos.shape)
    i = np.arange(n)
    a = (i[None, :] + (n + 1) * i[:, None]).reshape(-1)
    idx = np.concatenate([np.stack([a, a + 1, a + n + 2], 1), np.stack([a, a + n + 2, a + n + 1], 1)], 0)
    return pos, nrm, idx


def _merge(parts):
    pos, nrm, idx, off = [], [], [], 0
    for p, n, i in parts:
        pos.append(p); nrm.append(n); idx.append(i + off); off += p.shape[0]
    return (np.concatenate(pos).astype(np.float32), np.concatenate(nrm).astype(np.float32),
            np.concatenate(idx).astype(np.uint32))


def box_faces(dx, dy, dz, n=1, origin=(0.0, 0.0, 0.0)):
    """The 6 faces of OCCT's `box b dx dy dz` in `explode b FACE` order: -X +X -Y +Y -Z +Z."""
    o = np.array(origin, dtype=np.float64)
    X, Y, Z = np.array([dx, 0, 0.0]), np.array([0, dy, 0.0]), np.array([0, 0, dz])
    return [
        _grid_face(o, Z, Y, np.array([-1.0, 0, 0]), n),
        _grid_face(o + X, Y, Z, np.array([1.0, 0, 0]), n),
        _grid_face(o, X, Z, np.array([0, -1.0, 0]), n),
        _grid_face(o + Y, Z, X, np.array([0, 1.0, 0]), n),
        _grid_face(o, Y, X, np.array([0, 0, -1.0]), n),
        _grid_face(o + Z, X, Y, np.array([0, 0, 1.0]), n),
    ]


def box(dx, dy, dz, n=1, origin=(0.0, 0.0, 0.0)):
    return _merge(box_faces(dx, dy, dz, n, origin))


def l_bracket(a, b, t, n=1):
    """Two overlapping slabs forming an L profile."""
    return _merge(box_faces(a, t, b, n) + box_faces(t, a, b, n))


def uv_sphere(r, nu=64, nv=32, radii: Optional[np.ndarray] = None):
    """`psphere s r`: nu longitude x nv latitude segments, smooth normals; 2*nu*(nv-1) triangles."""
    th = np.linspace(0.0, math.pi, nv + 1)
    ph = np.linspace(0.0, 2.0 * math.pi, nu + 1)
    tt, pp = np.meshgrid(th, ph, indexing="ij")
    n = np.stack([np.sin(tt) * np.cos(pp), np.sin(tt) * np.sin(pp), np.cos(tt)], -1).reshape(-1, 3)
    rr = r if radii is None else (r * radii.reshape(-1, 1))
    pos = n * rr
    idx = []
    for i in range(nv):
        a = i * (nu + 1) + np.arange(nu)
        b = a + nu + 1
        if i > 0:
            idx.append(np.stack([a, b, a + 1], 1))
        if i < nv - 1:
            idx.append(np.stack([a + 1, b, b + 1], 1))
    return pos.astype(np.float32), n.astype(np.float32), np.concatenate(idx).astype(np.uint32)


def cylinder(r, h, seg=48, rings=1):
    ph = np.linspace(0.0, 2.0 * math.pi, seg + 1)
    z = np.linspace(0.0, h, rings + 1)
    zz, pp = np.meshgrid(z, ph, indexing="ij")
    n = np.stack([np.cos(pp), np.sin(pp), np.zeros_like(pp)], -1).reshape(-1, 3)
    pos = n * r + np.stack([np.zeros_like(zz), np.zeros_like(zz), zz], -1).reshape(-1, 3)
    idx = []
    for i in range(rings):
        a = i * (seg + 1) + np.arange(seg)
        b = a + seg + 1
        idx.append(np.stack([a, a + 1, b + 1], 1))
        idx.append(np.stack([a, b + 1, b], 1))
    side = (pos, n, np.concatenate(idx))
    caps = []
    for zc, nz in ((0.0, -1.0), (h, 1.0)):
        ring = np.stack([r * np.cos(ph[:-1]), r * np.sin(ph[:-1]), np.full(seg, zc)], 1)
        p = np.concatenate([np.array([[0.0, 0.0, zc]]), ring])
        k = np.arange(seg)
        tri = np.stack([np.zeros(seg, int), 1 + k, 1 + (k + 1) % seg], 1)
        if nz < 0:
            tri = tri[:, ::-1]
        caps.append((p, np.broadcast_to(np.array([0.0, 0.0, nz]), p.shape), tri))
    return _merge([side] + caps)


def torus(R, r, nu=48, nv=24):
    u = np.linspace(0.0, 2.0 * math.pi, nu + 1)
    v = np.linspace(0.0, 2.0 * math.pi, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    n = np.stack([np.cos(vv) * np.cos(uu), np.cos(vv) * np.sin(uu), np.sin(vv)], -1).reshape(-1, 3)
    c = np.stack([R * np.cos(uu), R * np.sin(uu), np.zeros_like(uu)], -1).reshape(-1, 3)
    pos = c + r * n
    idx = []
    for i in range(nu):
        a = i * (nv + 1) + np.arange(nv)
        b = a + nv + 1
        idx.append(np.stack([a, b, b + 1], 1))
        idx.append(np.stack([a, b + 1, a + 1], 1))
    return pos.astype(np.float32), n.astype(np.float32), np.concatenate(idx).astype(np.uint32)


# ------------------------------------------------------------------ transforms (row-major 3x4)

def trsf(translate=(0, 0, 0), rot_axis=(0, 0, 1), rot_deg=0.0, scale=1.0) -> np.ndarray:
    a = np.array(rot_axis, dtype=np.float64)
    a /= np.linalg.norm(a)
    t = math.radians(rot_deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    Rm = np.eye(3) + math.sin(t) * K + (1 - math.cos(t)) * (K @ K)
    m = np.zeros((3, 4))
    m[:, :3] = Rm * scale
    m[:, 3] = translate
    return m.astype(np.float32)


def random_rotation(rng: PCG32) -> np.ndarray:
    u1, u2, u3 = rng.uniform(), rng.uniform(), rng.uniform()
    q = np.array([math.sqrt(1 - u1) * math.sin(2 * math.pi * u2), math.sqrt(1 - u1) * math.cos(2 * math.pi * u2),
                  math.sqrt(u1) * math.sin(2 * math.pi * u3), math.sqrt(u1) * math.cos(2 * math.pi * u3)])
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# ------------------------------------------------------------------ scene description

@dataclass
class SceneDesc:
    name: str
    meshes: List[Tuple[np.ndarray, np.ndarray, np.ndarray]] = field(default_factory=list)
    instances: List[Tuple[int, Optional[np.ndarray], int]] = field(default_factory=list)  # mesh, xf, material
    materials: List[Graphic3d_BSDF] = field(default_factory=list)
    lights: list = field(default_factory=list)
    camera: Graphic3d_Camera = field(default_factory=Graphic3d_Camera)
    params: Graphic3d_RenderingParams = field(default_factory=Graphic3d_RenderingParams)
    width: int = 512
    height: int = 512
    envmap: Optional[np.ndarray] = None
    textures: List[np.ndarray] = field(default_factory=list)   # uint8 (h,w,3|4), referenced by Graphic3d_BSDF.TextureId
    mesh_uvs: dict = field(default_factory=dict)                # mesh index -> (n,2) float32 texel coordinates
    env_source: str = ""                                        # where the environment map came from (default_env)

    def add(self, mesh, xf=None, bsdf: Optional[Graphic3d_BSDF] = None, material_id: Optional[int] = None) -> int:
        self.meshes.append(mesh)
        if material_id is None:
            self.materials.append(bsdf if bsdf is not None else Graphic3d_BSDF.CreateDiffuse((0.8, 0.8, 0.8)))
            material_id = len(self.materials) - 1
        self.instances.append((len(self.meshes) - 1, xf, material_id))
        return len(self.instances) - 1

    def n_triangles(self) -> int:
        return int(sum(self.meshes[m][2].shape[0] for m, _, _ in self.instances))

    def apply(self, view: V3d_View, with_target: bool = True):
        """Feeds the scene through the host mirror (and so through the C-ABI)."""
        view.Clear()
        view.ClearTextures()
        for t in self.textures:
            view.AddTexture(t)
        ids = [view.AddMesh(p, i, n, self.mesh_uvs.get(k)) for k, (p, n, i) in enumerate(self.meshes)]
        for m, xf, mat in self.instances:
            view.Display(ids[m], xf, mat)
        view.SetMaterials(self.materials)
        view.SetLights(self.lights)
        view.SetTextureEnv(self.envmap)
        view.SetRenderingParams(self.params)
        self.camera.Aspect = self.width / self.height
        view.SetCamera(self.camera)
        if with_target:
            view.SetWindowSize(self.width, self.height)
        view.Update()


def look_at(eye, at, up=(0, 0, 1), fovy=45.0) -> Graphic3d_Camera:
    d = np.array(at, dtype=np.float64) - np.array(eye, dtype=np.float64)
    return Graphic3d_Camera(Eye=tuple(eye), Direction=tuple(d), Up=tuple(up), FOVy=fovy)


# ------------------------------------------------------------------ C1: Cornell box

def cornell_box(width=512, height=512, depth=4, sphere_res=(64, 32)) -> SceneDesc:
    """Mirror of data/scripts/CornellBox.tcl (OCCT prims tessellated here)."""
    s = SceneDesc("cornell", width=width, height=height)
    faces = box_faces(1, 1, 1)                        # CornellBox.tcl:20-21  box b 1 1 1; explode b FACE
    plaster = lambda kd: Graphic3d_BSDF(Kd=list(kd))  # :34-38  vbsdf b_i -kd ... -ks 0
    walls = [(0, (1, 0, 0), (1, 0.3, 0.3)),           # b_1 -> x = 1, red      (:23,34)
             (1, (-1, 0, 0), (0.3, 0.5, 1)),          # b_2 -> x = 0, blue     (:24,35)
             (2, (0, 1, 0), (1, 1, 1)),               # b_3 -> y = 1, back     (:25,36)
             (4, (0, 0, 1), (1, 1, 1)),               # b_5 -> z = 1, ceiling  (:26,37)
             (5, (0, 0, -1), (1, 1, 1))]              # b_6 -> z = 0, floor    (:27,38)
    for f, loc, kd in walls:
        s.add(_merge([faces[f]]), trsf(loc), plaster(kd))
    nu, nv = sphere_res
    # :44-49 glass sphere r 0.2 at (0.21,0.3,0.2), absorpColor .8 .8 1, absorpCoeff 6
    glass = Graphic3d_BSDF.CreateGlass((1, 1, 1), (0.8, 0.8, 1.0), 6.0, 1.5)
    s.add(uv_sphere(0.2, nu, nv), trsf((0.21, 0.3, 0.2)), glass)
    # :52-57 box c .3 .3 .2 at (0.55,0.3,0) rotated -30 deg about Z; -kd 1 .8 .2 -ks .3 -n
    yellow = Graphic3d_BSDF(Kd=[1.0, 0.8, 0.2], Ks=[0.3, 0.3, 0.3, 0.2],
                            FresnelBase=Graphic3d_Fresnel.CreateSchlick(0.04, 0.04, 0.04)).Normalize()
    s.add(box(0.3, 0.3, 0.2), trsf((0.55, 0.3, 0.0), (0, 0, 1), -30.0), yellow)
    # :60-66 glass box .15 .15 .3 at (0.7,0.25,0.2) rotated 10 deg; absorpColor .8 1 .8 coeff 6
    glass2 = Graphic3d_BSDF.CreateGlass((1, 1, 1), (0.8, 1.0, 0.8), 6.0, 1.5)
    s.add(box(0.15, 0.15, 0.3), trsf((0.7, 0.25, 0.2), (0, 0, 1), 10.0), glass2)
    # :69-74 sphere r 0.1 at (0.5,0.65,0.1): -kd .5 .9 .3 -ks .3 -baseRoughness 0 -n, baseFresnel Constant 1
    green = Graphic3d_BSDF(Kd=[0.5, 0.9, 0.3], Ks=[0.3, 0.3, 0.3, 0.0],
                           FresnelBase=Graphic3d_Fresnel.CreateConstant(1.0)).Normalize()
    s.add(uv_sphere(0.1, nu, nv), trsf((0.5, 0.65, 0.1)), green)
    # :12-14 positional light (0.5,0.5,0.85), smoothness (radius) 0.06, intensity 25
    s.lights = [make_light(True, (0.5, 0.5, 0.85), intensity=25.0, smoothness=0.06)]
    # :40-41 vfront; vfit with a perspective camera
    s.camera = look_at((0.5, -1.25, 0.5), (0.5, 0.5, 0.5), fovy=45.0)
    s.params = Graphic3d_RenderingParams(RaytracingDepth=depth)   # :76 uses -rayDepth 5; config C1 says 4
    return s


# ------------------------------------------------------------------ C2: STEP-like assembly

def _part_mesh(kind: int, res: int):
    if kind == 0:
        return box(1.0, 0.8, 0.6, n=max(1, res // 2), origin=(-0.5, -0.4, -0.3))
    if kind == 1:
        return cylinder(0.4, 1.0, seg=3 * res, rings=max(1, res // 2))
    if kind == 2:
        return torus(0.45, 0.18, nu=2 * res, nv=res)
    if kind == 3:
        return uv_sphere(0.5, nu=2 * res, nv=res)
    return l_bracket(0.9, 0.7, 0.2, n=max(1, res // 3))


def assembly(n_parts=1000, target_tris=1_000_000, seed=2, width=1920, height=1080, depth=8) -> SceneDesc:
    """Config C2: parts on a jittered cubic grid, each a randomly scaled/rotated primitive with
    its own mesh and material (60 % diffuse, 40 % glossy); ground quad; one directional light."""
    rng = PCG32(seed)
    side = max(1, round(n_parts ** (1.0 / 3.0)))
    kinds = [rng.randint(5) for _ in range(n_parts)]
    # resolution giving the requested total within 1 %
    def total(res):
        cache = {k: _part_mesh(k, res)[2].shape[0] for k in set(kinds)}
        return sum(cache[k] for k in kinds)
    res = 4
    while total(res + 1) <= target_tris and res < 256:
        res += 1
    base = {k: _part_mesh(k, res) for k in set(kinds)}
    fine = {k: _part_mesh(k, res + 1) for k in set(kinds)}
    cur = total(res)
    s = SceneDesc("assembly", width=width, height=height)
    cell = 1.6
    for i, k in enumerate(kinds):
        gx, gy, gz = i % side, (i // side) % side, i // (side * side)
        use_fine = cur + (fine[k][2].shape[0] - base[k][2].shape[0]) <= target_tris
        pos, nrm, idx = fine[k] if use_fine else base[k]
        if use_fine:
            cur += fine[k][2].shape[0] - base[k][2].shape[0]
        sc = np.array([rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0)])
        p = (pos * sc).astype(np.float32)
        n = nrm / sc
        n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
        m = np.zeros((3, 4), dtype=np.float32)
        m[:, :3] = random_rotation(rng)
        m[:, 3] = [(gx + 0.5 + rng.uniform(-0.25, 0.25)) * cell, (gy + 0.5 + rng.uniform(-0.25, 0.25)) * cell,
                   (gz + 0.5 + rng.uniform(-0.25, 0.25)) * cell + 0.3]
        if rng.uniform() < 0.6:
            bsdf = Graphic3d_BSDF(Kd=[rng.uniform(0.2, 0.9) for _ in range(3)])
        else:
            bsdf = Graphic3d_BSDF(Kd=[rng.uniform(0.1, 0.5) for _ in range(3)], Ks=[0.5, 0.5, 0.5, rng.uniform(0.02, 0.4)],
                                  FresnelBase=Graphic3d_Fresnel.CreateSchlick(0.04, 0.04, 0.04)).Normalize()
        s.add((p, n, idx), m, bsdf)
    ext = side * cell
    s.add(_merge([_grid_face(np.array([-ext, -ext, 0.0]), np.array([3 * ext, 0, 0.0]), np.array([0, 3 * ext, 0.0]),
                             np.array([0, 0, 1.0]), 1)]), None, Graphic3d_BSDF(Kd=[0.6, 0.6, 0.6]))
    # as data/scripts/Materials.tcl:203
    s.lights = [make_light(False, (-0.303949, -0.434084, -0.848048), intensity=12.0, smoothness=0.3)]
    c = ext * 0.5
    s.camera = look_at((c + ext * 1.15, c - ext * 1.35, c + ext * 0.95), (c, c, c * 0.8), fovy=40.0)
    s.params = Graphic3d_RenderingParams(RaytracingDepth=depth)
    return s


# ------------------------------------------------------------------ C3: Materials.tcl

def materials_scene(width=1920, height=1080, depth=12, sphere_res=(128, 64), env: Optional[np.ndarray] = None) -> SceneDesc:
    """Mirror of data/scripts/Materials.tcl: 9 balls r 10 on a 12x12 chess floor."""
    s = SceneDesc("materials", width=width, height=height)
    tile = box(10, 10, 0.1)                                           # Materials.tcl:21
    s.materials = []
    light_tile = Graphic3d_BSDF(Kd=[0.85] * 3)                        # :31
    dark_tile = Graphic3d_BSDF(Kd=[0.45] * 3)                         # :33
    s.materials += [light_tile, dark_tile]
    s.meshes.append(tile)
    for i in range(12):                                               # :24-36
        for j in range(1, 13):
            xf = trsf((i * 10 - 90, j * 10 - 70, -0.15))
            s.instances.append((0, xf, 0 if (i + j) % 2 == 0 else 1))
    F = Graphic3d_Fresnel
    brass = F.CreateSchlick(0.58, 0.42, 0.2)
    balls = [
        # Ball1 :40-54
        ((10, 0, 10), Graphic3d_BSDF(Kd=[0.272798, 0.746262, 0.104794], Ks=[0.253738] * 3 + [0.045],
                                     FresnelCoat=F.CreateConstant(0), FresnelBase=brass)),
        # Ball2 :57-71 (emissive)
        ((10, 40, 10), Graphic3d_BSDF(Kd=[0.8] * 3, Le=[2.02, 0.171915, 0.171915])),
        # Ball3 :74-88 glass, absorption .75 .95 .9 / 0.05
        ((-30, -40, 10), Graphic3d_BSDF(Kc=[1, 1, 1, 0], Kt=[1, 1, 1], Absorption=[0.75, 0.95, 0.9, 0.05],
                                        FresnelCoat=F.CreateDielectric(1.62))),
        # Ball4 :91-105 mirror metal
        ((-70, -40, 10), Graphic3d_BSDF(Ks=[0.985] * 3 + [0.0], FresnelBase=brass)),
        # Ball5 :108-122 blue glass
        ((-30, 0, 10), Graphic3d_BSDF(Kc=[1, 1, 1, 0], Kt=[1, 1, 1], Absorption=[0, 0.288061, 0.825532, 0.3],
                                      FresnelCoat=F.CreateDielectric(1.62))),
        # Ball6 :125-139 car paint: dielectric coat over diffuse + glossy
        ((-30, 40, 10), Graphic3d_BSDF(Kc=[1, 1, 1, 0], Kd=[0, 0.716033, 0.884507], Ks=[0.115493] * 3 + [0.045],
                                       FresnelCoat=F.CreateDielectric(1.5), FresnelBase=brass)),
        # Ball7 :142-156 coated rough metal
        ((-70, 0, 10), Graphic3d_BSDF(Kc=[1, 1, 1, 0], Kd=[1e-06, 9.9999e-07, 9.9999e-07], Ks=[0.0479573, 0.804998, 0, 0.447],
                                      FresnelCoat=F.CreateDielectric(1.5), FresnelBase=brass)),
        # Ball8 :159-173 aluminium
        ((-70, 40, 10), Graphic3d_BSDF(Ks=[0.985] * 3 + [0.026], FresnelBase=F.CreateSchlick(0.913183, 0.921494, 0.924524))),
        # Ball0 :176-190 red diffuse
        ((10, -40, 10), Graphic3d_BSDF(Kd=[0.723404, 0.166229, 0.166229])),
    ]
    nu, nv = sphere_res
    ball = uv_sphere(10.0, nu, nv)                                    # :10-18 psphere BallN 10
    s.meshes.append(ball)
    for loc, b in balls:
        s.materials.append(b)
        s.instances.append((1, trsf(loc), len(s.materials) - 1))
    # :193-199 camera; :202-203 light
    s.camera = Graphic3d_Camera(Eye=(139.412, -1.62643, 178.037),
                                Direction=(-22.3025 - 139.412, 0.0986351 + 1.62643, 3.30327 - 178.037),
                                Up=(-0.733931, -0.00311795, 0.679217), FOVy=25.0)
    s.lights = [] if env is not None else [make_light(False, (-0.303949, -0.434084, -0.848048), intensity=12.0, smoothness=0.3)]
    s.envmap = env
    s.params = Graphic3d_RenderingParams(RaytracingDepth=depth)
    return s


def synthetic_env(width=2048, height=1024) -> np.ndarray:
    """Procedural lat-long sky of the shape of data/maps/default.jpg (2048x1024; the JPEG itself is the
    reference's data and is not shipped): horizon-to-zenith gradient, a warm sun with a halo, dark ground.
    Float radiance, deterministic."""
    v = (np.arange(height, dtype=np.float32) + 0.5) / height          # 0 = up
    u = (np.arange(width, dtype=np.float32) + 0.5) / width
    el = (0.5 - v)[:, None] * np.float32(np.pi)                        # elevation
    az = (u[None, :] - 0.5) * np.float32(2 * np.pi)
    up = np.clip(np.sin(el), 0, 1)
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.50 + 0.20 * (1 - up), 0.85 - 0.15 * (1 - up)], axis=-1) * (0.6 + 0.6 * up[..., None])
    sky = np.broadcast_to(sky, (height, width, 3)).copy()
    ground = np.array([0.18, 0.16, 0.14], np.float32)
    img = np.where((el > 0)[..., None], sky, ground[None, None, :]).astype(np.float32)
    sun_el, sun_az = np.float32(0.75), np.float32(-0.9)
    d = np.sin(el) * np.sin(sun_el) + np.cos(el) * np.cos(sun_el) * np.cos(az - sun_az)
    ang = np.arccos(np.clip(d, -1, 1))
    img += (np.exp(-(ang / 0.035) ** 2) * 60.0 + np.exp(-(ang / 0.25) ** 2) * 0.8)[..., None] * np.array([1.0, 0.93, 0.8], np.float32)
    return np.ascontiguousarray(img, dtype=np.float32)


def default_env(path: Optional[str] = None):
    """The environment map CADRays loads by default (data/maps/default.jpg, AppGui.cxx:963), 2048x1024, as
    8-bit RGB (the library linearises 8-bit texels as (c/255)^2).  Looked up in this order: `path`,
    $CADRAYS_DATA_DIR/maps/default.jpg, the reference checkout, the decoded copy of the same pixels under
    tests/golden/ (made by tests/golden/make_default_env.py; the one that exists on a GPU box).  Returns
    (image or None, where it came from); None = not found, callers fall back to synthetic_env()."""
    import os
    from pathlib import Path
    here = Path(__file__).resolve().parent.parent
    cands = []
    if path:
        cands.append((Path(path), str(path)))
    if os.environ.get("CADRAYS_DATA_DIR"):
        cands.append((Path(os.environ["CADRAYS_DATA_DIR"]) / "maps" / "default.jpg", "$CADRAYS_DATA_DIR/maps/default.jpg"))
    cands.append((Path("/root/reference/data/maps/default.jpg"), "reference data/maps/default.jpg"))
    cands.append((here / "tests" / "golden" / "default_env_2048x1024.png",
                  "data/maps/default.jpg (decoded copy: tests/golden/default_env_2048x1024.png)"))
    for p, what in cands:
        if not p.exists():
            continue
        try:
            from PIL import Image
            a = np.asarray(Image.open(p).convert("RGB"))
        except ImportError:
            if p.suffix.lower() != ".png":
                continue
            from .imageio import read_png_rgb8
            a = read_png_rgb8(str(p))
        return np.ascontiguousarray(a, dtype=np.uint8), what
    return None, "not found"


def product_shot(width=3840, height=2160, depth=12, sphere_res=(128, 64), env_path: Optional[str] = None) -> SceneDesc:
    """Config C4: the Materials.tcl geometry lit only by data/maps/default.jpg (2048x1024 lat-long), 3840x2160.
    s.env_source says which file the map came from; the procedural sky is the fallback when none is found."""
    env, src = default_env(env_path)
    if env is None:
        env, src = synthetic_env(), "synthetic sky (data/maps/default.jpg not found)"
    s = materials_scene(width, height, depth, sphere_res, env=env)
    s.name = "product_shot"
    s.env_source = src
    return s


# ------------------------------------------------------------------ C5: instanced stress

def instanced(n_inst=1024, n_meshes=16, seed=5, width=1920, height=1080, depth=8, nu=80, nv=65) -> SceneDesc:
    """Config C5: n_inst instances of n_meshes bumpy spheres (2*nu*(nv-1) = 10 240 triangles each)."""
    rng = PCG32(seed)
    s = SceneDesc("instanced", width=width, height=height)
    for m in range(n_meshes):
        th = np.linspace(0.0, math.pi, nv + 1)[:, None]
        ph = np.linspace(0.0, 2.0 * math.pi, nu + 1)[None, :]
        k1, k2 = 2 + rng.randint(5), 1 + rng.randint(4)
        a1, a2 = rng.uniform(0.03, 0.12), rng.uniform(0.03, 0.1)
        radii = 1.0 + a1 * np.sin(k1 * ph) * np.sin(th) ** 2 + a2 * np.cos(k2 * 2 * th) * np.ones_like(ph)
        radii[:, -1] = radii[:, 0]
        pos, _, idx = uv_sphere(0.5, nu, nv, radii)
        s.meshes.append((pos, None, idx))
    n_mat = 32
    for k in range(n_mat):
        if k % 2 == 0:
            s.materials.append(Graphic3d_BSDF(Kd=[rng.uniform(0.2, 0.9) for _ in range(3)]))
        else:
            s.materials.append(Graphic3d_BSDF(Kd=[rng.uniform(0.1, 0.4) for _ in range(3)], Ks=[0.5, 0.5, 0.5, rng.uniform(0.05, 0.3)],
                                              FresnelBase=Graphic3d_Fresnel.CreateSchlick(0.04, 0.04, 0.04)).Normalize())
    side = max(1, math.ceil(math.sqrt(n_inst / 4.0)))
    for i in range(n_inst):
        gx, gy, gz = i % side, (i // side) % side, i // (side * side)
        m = np.zeros((3, 4), dtype=np.float32)
        m[:, :3] = random_rotation(rng) * rng.uniform(0.7, 1.2)
        m[:, 3] = [gx * 1.3 + rng.uniform(-0.2, 0.2), gy * 1.3 + rng.uniform(-0.2, 0.2), 0.6 + gz * 1.3 + rng.uniform(-0.1, 0.1)]
        s.instances.append((i % n_meshes, m, rng.randint(n_mat)))
    ext = side * 1.3
    s.materials.append(Graphic3d_BSDF(Kd=[0.55, 0.55, 0.55]))
    s.meshes.append(_merge([_grid_face(np.array([-ext, -ext, 0.0]), np.array([3 * ext, 0, 0.0]), np.array([0, 3 * ext, 0.0]),
                                       np.array([0, 0, 1.0]), 1)]))
    s.instances.append((len(s.meshes) - 1, None, len(s.materials) - 1))
    s.lights = [make_light(False, (-0.303949, -0.434084, -0.848048), intensity=12.0, smoothness=0.3)]
    c = ext * 0.5
    s.camera = look_at((c + ext * 0.9, c - ext * 1.2, ext * 0.9), (c, c, 1.5), fovy=40.0)
    s.params = Graphic3d_RenderingParams(RaytracingDepth=depth)
    return s


def random_rays(n: int, lo, hi, seed: int = 7):
    """Incoherent test rays: origins uniform in the box [lo,hi] grown by 50 %, directions uniform on the sphere."""
    g = np.random.default_rng(seed)
    lo = np.asarray(lo, dtype=np.float64); hi = np.asarray(hi, dtype=np.float64)
    c, e = 0.5 * (lo + hi), 0.75 * (hi - lo)
    org = (c + (g.random((n, 3)) * 2 - 1) * e).astype(np.float32)
    d = g.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return org, d.astype(np.float32)


def deep_tree_scene(n=700, base=1.1, n_inst=100, inst_base=1.15, width=64, height=48) -> SceneDesc:
    """Degenerate test scene: exponentially spaced triangles / instances make the SAH builder peel a few
    primitives per level, so the two-level tree is deeper (> 28 levels together) than the traversal's
    shared-memory stack and exercises its local-memory overflow."""
    d = SceneDesc("deep", width=width, height=height)
    k = np.arange(n)
    x = base ** k * 1e-3
    s = x * 0.04
    z = np.zeros(n)
    pos = np.stack([np.stack([x - s, -s, z], 1), np.stack([x + s, -s, z], 1), np.stack([x, s, z], 1)], 1).reshape(-1, 3)
    d.meshes.append((pos.astype(np.float32), np.tile(np.array([[0, 0, 1]], np.float32), (3 * n, 1)),
                     np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)))
    d.materials.append(Graphic3d_BSDF(Kd=[0.7, 0.7, 0.7]))
    for j in range(n_inst):
        d.instances.append((0, trsf((0.0, 1e-2 * inst_base ** j, 0.0)), 0))
    d.lights = [make_light(False, (0.1, 0.2, -1.0), intensity=3.0, smoothness=0.1)]
    d.camera = look_at((20.0, 20.0, 120.0), (20.0, 20.0, 0.0), up=(0, 1, 0), fovy=40.0)
    d.params = Graphic3d_RenderingParams(RaytracingDepth=3)
    return d
