"""CPU tests of the product's host side (no GPU, no compute calls): the C-ABI library loads and
exports every symbol include/cadrays_b200.h declares, the host BVH builder produces a well-formed
two-level tree, the host mirror keeps CADRays' material semantics, and device entry points fail
loudly without a device."""
import os
import re
import struct
from pathlib import Path

import numpy as np
import pytest

from cadrays_b200 import _ffi, scenes
from cadrays_b200._ffi import CRT_ERR_INVALID_ARG, CRT_ERR_NO_DEVICE, CRT_ERR_STATE, CrtError
from cadrays_b200.view import (Graphic3d_BSDF, Graphic3d_Fresnel, Graphic3d_RenderingParams, V3d_View)

REPO = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(product_lib):
    header = (REPO / "include" / "cadrays_b200.h").read_text()
    declared = set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", header))
    declared -= {"crt_bsdf", "crt_light", "crt_params", "crt_camera", "crt_stats", "crt_context", "crt_status"}
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(product_lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_ffi.PROTOTYPES), "ctypes prototypes and header disagree"
    version = int(re.search(r"#define CRT_ABI_VERSION (\d+)", header).group(1))
    assert product_lib.crt_abi_version() == version == 3


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_ffi.crt_bsdf) == 128
    assert C.sizeof(_ffi.crt_light) == 32
    assert C.sizeof(_ffi.crt_stats) == 112
    assert C.sizeof(_ffi.crt_camera) == 52
    assert C.sizeof(_ffi.crt_params) == 76


def test_no_device_fails_loudly(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(CrtError) as e:
        V3d_View(0)
    assert e.value.code == CRT_ERR_NO_DEVICE
    assert "no CPU fallback" in e.value.message
    v = V3d_View(host_only=True)
    scenes.cornell_box(32, 32, sphere_res=(8, 4)).apply(v, with_target=False)
    for call in (lambda: v.Redraw(1), lambda: v.SetWindowSize(8, 8),
                 lambda: v.Trace(np.zeros((1, 3), np.float32), np.ones((1, 3), np.float32)),
                 lambda: v.ImportBVH(v.ExportBVH())):
        with pytest.raises(CrtError) as e:
            call()
        assert e.value.code in (CRT_ERR_NO_DEVICE, CRT_ERR_STATE)
    v.Remove()


def test_product_never_imports_the_oracle():
    """No include, import, link or dlopen of anything under oracle/ from the product tree (the one dlopen the
    product makes is NCCL's own library, for crt_group's optional ncclReduce path)."""
    bad = re.compile(r"(#\s*include[^\n]*oracle|^\s*(from|import)\s+oracle|oracle_ffi|libcadrays_oracle|dlopen\s*\((?!\"libnccl\.so\.2\"))", re.M)
    files = list((REPO / "cadrays_b200").rglob("*.py")) + [p for p in (REPO / "cadrays_b200" / "csrc").glob("*") if p.is_file()]
    files.append(REPO / "include" / "cadrays_b200.h")
    files.append(REPO / "include" / "cadrays_b200.hpp")
    files += [p for p in (REPO / "tools").glob("*") if p.is_file()]      # measurement tooling drives the product only
    for p in files:
        m = bad.search(p.read_text(errors="ignore"))
        assert m is None, (p, m.group(0))
    from cadrays_b200 import build
    assert not any("oracle" in str(x) for x in build._sources())
    # bench.py and __graft_entry__.py may use the oracle, but only in the CPU legs / smoke(): never at import time
    for name in ("bench.py", "__graft_entry__.py"):
        top_level = [ln for ln in (REPO / name).read_text().splitlines() if re.match(r"^(from|import)\s+oracle", ln)]
        assert not top_level, (name, top_level)


def _parse_blob(blob):
    h = struct.unpack_from("<8I7fI", blob, 0)
    hdr = dict(zip(("magic", "version", "n_nodes", "n_verts", "n_tris", "n_inst", "n_top", "flags"), h[:8]))
    hdr["min"], hdr["max"], hdr["eps"] = np.array(h[8:11]), np.array(h[11:14]), h[14]
    off = 64
    def take(dtype, count, width):
        nonlocal off
        a = np.frombuffer(blob, dtype=dtype, count=count * width, offset=off).reshape(count, width)
        off = (off + a.nbytes + 15) & ~15
        return a
    n, v, t, i = hdr["n_nodes"], hdr["n_verts"], hdr["n_tris"], hdr["n_inst"]
    out = dict(hdr=hdr, info=take(np.int32, n, 4), bmin=take(np.float32, n, 3), bmax=take(np.float32, n, 3),
               pos=take(np.float32, v, 3), nrm=take(np.float32, v, 3), uv=take(np.float32, v, 2),
               tris=take(np.int32, t, 4), inv=take(np.float32, i, 16), meta=take(np.int32, i, 4))
    assert off <= len(blob)
    return out


@pytest.mark.parametrize("which", ["cornell", "assembly", "instanced"])
def test_bvh_blob_is_well_formed(which, product_lib):
    desc = {"cornell": lambda: scenes.cornell_box(32, 32, sphere_res=(24, 12)),
            "assembly": lambda: scenes.assembly(n_parts=40, target_tris=20000, width=32, height=32),
            "instanced": lambda: scenes.instanced(n_inst=30, n_meshes=3, width=32, height=32, nu=16, nv=9)}[which]()
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    b = _parse_blob(v.ExportBVH())
    v.Remove()
    hdr = b["hdr"]
    assert hdr["magic"] == 0x42545243 and hdr["version"] == 1 and hdr["n_inst"] == len(desc.instances)
    assert hdr["eps"] == pytest.approx(max(1e-6, 1e-4 * float(np.linalg.norm(hdr["max"] - hdr["min"]))), rel=1e-5)
    info, bmin, bmax = b["info"], b["bmin"], b["bmax"]
    seen_inst = set()
    # top level: leaf size 1, every instance exactly once, child boxes inside the parent's
    stack = [(0, 0)]
    while stack:
        n, depth = stack.pop()
        assert depth <= 32
        x, y, z, w = info[n]
        if x == 0:
            for c in (y, z):
                assert (bmin[c] >= bmin[n] - 1e-6).all() and (bmax[c] <= bmax[n] + 1e-6).all()
                stack.append((c, depth + 1))
        else:
            assert x > 0 and x - 1 not in seen_inst
            seen_inst.add(x - 1)
            assert b["meta"][x - 1][2] == y
            # the (tight, padded) world box of the instance contains every transformed vertex
            m, xf, _ = desc.instances[x - 1]
            M = np.eye(3, 4) if xf is None else np.asarray(xf, np.float64)
            wpos = desc.meshes[m][0].astype(np.float64) @ M[:, :3].T + M[:, 3]
            assert (wpos >= bmin[n] - 1e-7).all() and (wpos <= bmax[n] + 1e-7).all()
            ext = wpos.max(0) - wpos.min(0)
            assert ((bmax[n] - bmin[n]) <= ext * 1.001 + 1e-4).all()      # and is tight
    assert seen_inst == set(range(hdr["n_inst"]))
    # bottom level, per distinct mesh: leaves partition the triangle range, leaf size <= 5,
    # every triangle inside its leaf box, children inside parents
    for root, voff, toff in {(r[1], r[2], r[3]) for r in info[:hdr["n_top"]] if r[0] > 0}:
        covered = []
        stack = [(root, 0)]
        while stack:
            n, depth = stack.pop()
            assert depth <= 32
            x, y, z, w = info[n]
            if x == 0:
                for c in (root + y, root + z):
                    assert (bmin[c] >= bmin[n] - 1e-6).all() and (bmax[c] <= bmax[n] + 1e-6).all()
                    stack.append((c, depth + 1))
            else:
                assert x < 0 and 0 < z - y + 1 <= 5
                for k in range(y, z + 1):
                    tri = b["tris"][toff + k]
                    p = b["pos"][voff + tri[:3]]
                    assert (p >= bmin[n] - 1e-6).all() and (p <= bmax[n] + 1e-6).all()
                    covered.append(k)
        covered.sort()
        assert covered == list(range(len(covered)))
    # inverse matrices really invert the instance transforms
    for k, (m, xf, mat) in enumerate(desc.instances):
        M = np.eye(4); M[:3] = np.eye(3, 4) if xf is None else xf
        Minv = b["inv"][k].reshape(4, 4)
        assert np.allclose(Minv @ M, np.eye(4), atol=1e-4)
        assert b["meta"][k][0] == mat


def test_builder_rejects_bad_input(product_lib):
    v = V3d_View(host_only=True)
    with pytest.raises(CrtError) as e:
        v.AddMesh(np.zeros((3, 3), np.float32), np.array([[0, 1, 3]], np.uint32))
    assert e.value.code == CRT_ERR_INVALID_ARG
    with pytest.raises(CrtError):
        v.AddMesh(np.array([[0, 0, np.inf], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    m = v.AddMesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    v.Display(m, np.zeros((3, 4), np.float32))          # singular transform
    with pytest.raises(CrtError) as e:
        v.Update()
    assert e.value.code == CRT_ERR_INVALID_ARG
    with pytest.raises(CrtError) as e:
        V3d_View(host_only=True).ExportBVH()              # export before commit
    assert e.value.code == CRT_ERR_STATE
    v.Remove()


def test_empty_and_single_triangle_scenes(product_lib, oracle_lib):
    from oracle.oracle_ffi import OracleScene
    v = V3d_View(host_only=True)
    v.Update()
    b = _parse_blob(v.ExportBVH())
    assert b["hdr"]["n_nodes"] == 0 and b["hdr"]["n_inst"] == 0
    o = OracleScene(v.ExportBVH())
    assert o.trace(np.zeros((2, 3), np.float32), np.ones((2, 3), np.float32))[0].tolist() == [-1, -1]
    m = v.AddMesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    v.Display(m)
    v.Update()
    b = _parse_blob(v.ExportBVH())
    assert b["hdr"]["n_nodes"] == 2 and b["info"][0][0] == 1 and b["info"][1][0] == -1   # top leaf + bottom leaf
    # vertex normals synthesised from the face when the caller gives none
    assert np.allclose(b["nrm"], [[0, 0, 1]] * 3)
    v.Remove()


def test_material_semantics_follow_cadrays():
    # MaterialEditor.cxx:294-329: clamp to [0,1], Le and absorption coefficient >= 0, normalise Kd+Ks+Kt
    b = Graphic3d_BSDF(Kd=[1.0, 0.8, 0.2], Ks=[0.3, 0.3, 0.3, 0.1], Le=[-1, 2, 3], Absorption=[2, -1, 0.5, -4]).Normalize()
    assert max(b.Kd[k] + b.Ks[k] + b.Kt[k] for k in range(3)) == pytest.approx(1.0)
    assert b.Kd[0] == pytest.approx(1.0 / 1.3) and b.Ks[3] == 0.1
    assert b.Le == [0.0, 2.0, 3.0] and b.Absorption == [1.0, 0.0, 0.5, 0.0]
    # Graphic3d_Fresnel::Serialize (MaterialEditor.cxx:209-255; ImportExport.cxx:204-227)
    assert Graphic3d_Fresnel.CreateSchlick(0.58, 0.42, 0.2).Serialize()[:3] == pytest.approx((0.58, 0.42, 0.2))
    assert Graphic3d_Fresnel.CreateConstant(0.7).Serialize()[0] == -1 and Graphic3d_Fresnel.CreateConstant(0.7).Serialize()[2] == pytest.approx(0.7)
    assert Graphic3d_Fresnel.CreateConductor(0.8, 5.8).Serialize()[:3] == pytest.approx((-2, 0.8, 5.8))
    assert Graphic3d_Fresnel.CreateDielectric(1.5).Serialize()[:2] == pytest.approx((-3, 1.5))
    g = Graphic3d_BSDF.CreateGlass((1, 1, 1), (0.8, 0.8, 1.0), 6.0, 1.5).to_c()
    assert list(g.Kt)[:3] == [1, 1, 1] and g.FresnelCoat[0] == -3 and g.Absorption[3] == 6.0
    with pytest.raises(ValueError):
        Graphic3d_RenderingParams(IsGlobalIlluminationEnabled=False).to_c()
    r = Graphic3d_RenderingParams(AdaptiveScreenSampling=True, NbRayTracingTiles=128).to_c()   # SettingsWidget.cxx:70-72
    assert r.adaptive_sampling == 1 and r.adaptive_tiles == 128


def test_scene_generators_are_seeded_and_sized():
    a = scenes.assembly(n_parts=30, target_tris=15000, seed=2)
    b = scenes.assembly(n_parts=30, target_tris=15000, seed=2)
    c = scenes.assembly(n_parts=30, target_tris=15000, seed=3)
    assert a.n_triangles() == b.n_triangles() and abs(a.n_triangles() - 15000) <= 0.05 * 15000
    assert all(np.array_equal(x[1], y[1]) for x, y in zip(a.instances, b.instances))
    assert not all(np.array_equal(x[1], y[1]) for x, y in zip(a.instances[:-1], c.instances[:-1]))
    inst = scenes.instanced(n_inst=8, n_meshes=2, nu=80, nv=65)
    assert inst.meshes[0][2].shape[0] == 10240                      # config C5 mesh size
    cb = scenes.cornell_box()
    assert len(cb.instances) == 9 and cb.lights[0].is_point == 1 and cb.lights[0].smoothness == pytest.approx(0.06)
    ms = scenes.materials_scene(sphere_res=(16, 8))
    assert len(ms.instances) == 144 + 9 and ms.camera.FOVy == 25.0


def test_cpp_host_mirror_compiles_and_runs(tmp_path, product_lib):
    """include/cadrays_b200.hpp (OCCT-named C++ wrapper) against the C-ABI: g++ build + host-only run."""
    import subprocess
    exe = tmp_path / "host_mirror_check"
    lib = REPO / "cadrays_b200" / "libcadrays_b200.so"
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", f"-I{REPO / 'include'}", str(REPO / "tests" / "cpp" / "host_mirror_check.cpp"),
           "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "blob" in r.stdout


def test_c_abi_header_is_plain_c(tmp_path, product_lib):
    """include/cadrays_b200.h compiles as C99 (-pedantic -Werror) and a C host can assemble a scene, set the
    parameters, build / export the BVH and gets a loud refusal from every device call without a GPU."""
    import subprocess
    exe = tmp_path / "c_abi_check"
    lib = REPO / "cadrays_b200" / "libcadrays_b200.so"
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{REPO / 'include'}", str(REPO / "tests" / "cpp" / "c_abi_check.c"),
           "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "c abi ok" in r.stdout


def test_blob_patch_shortcut_never_keeps_stale_geometry(product_lib):
    """A commit after instance-only edits rewrites just the header, the top-level nodes and the instance records of
    the previous blob.  It must not survive anything that changes geometry: Clear() followed by a different mesh
    with the same counts, another instance -> mesh map, or a change of tree width."""
    def fresh(meshes, inst, width=2):
        v = V3d_View(host_only=True)
        p = Graphic3d_RenderingParams(BvhWidth=width)
        v.SetRenderingParams(p)
        ids = [v.AddMesh(*m) for m in meshes]
        for m, xf in inst:
            v.Display(ids[m], xf, 0)
        v.Update()
        b = v.ExportBVH()
        v.Remove()
        return b

    pa, na, ia = scenes.uv_sphere(1.0, 24, 12)
    pb = (pa * np.array([1.0, 2.0, 0.5], np.float32)).astype(np.float32)          # same counts, other shape
    sphere_a, sphere_b = (pa, ia, na), (pb, ia, na)
    move = scenes.trsf((2, 0, 0))
    v = V3d_View(host_only=True)
    a = v.AddMesh(*sphere_a)
    v.Display(a, None, 0); v.Display(a, move, 0)
    v.Update()
    assert v.ExportBVH() == fresh([sphere_a], [(0, None), (0, move)])
    v.SetLocation(1, scenes.trsf((0, 3, 0)))                                        # shortcut applies
    v.Update()
    assert v.ExportBVH() == fresh([sphere_a], [(0, None), (0, scenes.trsf((0, 3, 0)))])
    v.Clear()                                                                       # same ids and sizes, other vertices
    b = v.AddMesh(*sphere_b)
    v.Display(b, None, 0); v.Display(b, scenes.trsf((0, 3, 0)), 0)
    v.Update()
    assert v.ExportBVH() == fresh([sphere_b], [(0, None), (0, scenes.trsf((0, 3, 0)))])
    c = v.AddMesh(*sphere_a)                                                        # second mesh, map changes
    v.Display(c, move, 0)
    v.Update()
    assert v.ExportBVH() == fresh([sphere_b, sphere_a], [(0, None), (0, scenes.trsf((0, 3, 0))), (1, move)])
    p = v.ChangeRenderingParams(); p.BvhWidth = 4                                   # tree width changes
    v.SetRenderingParams(p); v.Update()
    assert v.ExportBVH() == fresh([sphere_b, sphere_a], [(0, None), (0, scenes.trsf((0, 3, 0))), (1, move)], width=4)
    v.Remove()


def test_hidden_instances_leave_the_top_level_tree(product_lib, oracle_lib):
    """crt_instance_set_visible (AIS Erase / Display): a hidden instance keeps its id and records but no top-level
    leaf refers to it, so the oracle renders exactly the scene without that object; showing it again restores the
    original blob; hiding everything gives an empty scene."""
    from oracle.oracle_ffi import OracleScene
    desc = scenes.cornell_box(40, 40, depth=3, sphere_res=(12, 6))
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    original = v.ExportBVH()
    hidden = 6                                           # the yellow box
    v.SetVisible(hidden, False)
    v.Update()
    blob = v.ExportBVH()
    h = _parse_blob(blob)
    assert h["hdr"]["n_inst"] == len(desc.instances)
    leaves = h["info"][:h["hdr"]["n_top"], 0]
    assert hidden + 1 not in leaves and sorted(x for x in leaves if x > 0) == [k + 1 for k in range(len(desc.instances)) if k != hidden]
    o = OracleScene(blob); o.configure(desc)
    without = scenes.cornell_box(40, 40, depth=3, sphere_res=(12, 6))
    del without.instances[hidden]
    v2 = V3d_View(host_only=True)
    without.apply(v2, with_target=False)
    o2 = OracleScene(v2.ExportBVH()); o2.configure(without)
    assert np.array_equal(o.render(40, 40, 3), o2.render(40, 40, 3))
    o.close(); o2.close(); v2.Remove()
    v.SetVisible(hidden, True)
    v.Update()
    assert v.ExportBVH() == original
    for k in range(len(desc.instances)):
        v.SetVisible(k, False)
    v.Update()
    assert _parse_blob(v.ExportBVH())["hdr"]["n_inst"] == 0
    with pytest.raises(CrtError):
        v.SetVisible(999, True)
    v.Remove()


def test_parallel_tree_build_is_deterministic(monkeypatch, product_lib):
    """Meshes of 200 k triangles and more are built on several threads (upper levels with parallel passes over the
    wide nodes, sub-trees as independent tasks, then renumbered): the blob is byte-identical to the single-thread
    build, whatever the thread count."""
    pos, nrm, idx = scenes.uv_sphere(1.0, 480, 240)
    assert idx.shape[0] >= 200_000
    g = np.random.default_rng(3)
    pos = (pos + g.normal(scale=2e-3, size=pos.shape)).astype(np.float32)

    def blob():
        v = V3d_View(host_only=True)
        m = v.AddMesh(pos, idx, nrm)
        v.Display(m, None, 0)
        v.Display(m, scenes.trsf((3, 0, 0), (0, 1, 0), 30.0), 0)
        v.Update()
        b = v.ExportBVH()
        v.Remove()
        return b

    parallel = blob()
    monkeypatch.setenv("CRT_BUILD_SERIAL", "1")
    serial = blob()
    assert parallel == serial
    monkeypatch.delenv("CRT_BUILD_SERIAL")
    monkeypatch.setenv("OMP_NUM_THREADS", "3")      # read by libgomp at load time only; the std::thread passes still vary
    assert blob() == serial


def test_host_scene_under_sanitizers(tmp_path):
    """cadrays_b200/csrc/host_scene.cpp built with AddressSanitizer + UBSan and driven by tests/cpp/host_scene_sanitize.cpp:
    random scenes through build -> parse -> device layout for both tree widths, edit sequences (patched blob = fresh
    build), bit-flipped and truncated blobs, singular transforms."""
    import subprocess
    exe = tmp_path / "host_scene_sanitize"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fopenmp", "-Wall",
           str(REPO / "tests" / "cpp" / "host_scene_sanitize.cpp"), str(REPO / "cadrays_b200" / "csrc" / "host_scene.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "host scene sanitize ok" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])


def test_recommit_after_edit_equals_fresh_build(product_lib):
    """Scene edits (SetLocation / material index) reuse the cached bottom trees; the resulting blob is the one a
    fresh build of the edited scene gives, and undoing the edit restores the original bytes."""
    desc = scenes.assembly(n_parts=30, target_tris=9000, width=32, height=32)
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    original = v.ExportBVH()
    moved = scenes.trsf((1, 2, 3), (0, 1, 1), 40.0)
    v.SetLocation(4, moved)
    v.SetMaterialIndex(7, 2)
    v.Update()
    edited = v.ExportBVH()
    assert edited != original
    m, _, mat = desc.instances[4]
    desc.instances[4] = (m, moved, mat)
    m7, x7, _ = desc.instances[7]
    desc.instances[7] = (m7, x7, 2)
    w = V3d_View(host_only=True)
    desc.apply(w, with_target=False)
    assert w.ExportBVH() == edited
    v.SetLocation(4, scenes.assembly(n_parts=30, target_tris=9000, width=32, height=32).instances[4][1])
    v.SetMaterialIndex(7, mat if False else scenes.assembly(n_parts=30, target_tris=9000).instances[7][2])
    v.Update()
    assert v.ExportBVH() == original
    v.Remove(); w.Remove()


def test_quad_bvh_blob(product_lib, oracle_lib):
    """bvh_width = 4 (OCCT's optional QUAD_BVH collapse): inner nodes have 2..4 contiguous children that are
    the binary tree's grandchildren; the oracle finds the same hits through it as through the binary tree."""
    from oracle.oracle_ffi import OracleScene
    desc = scenes.assembly(n_parts=40, target_tris=20000, width=32, height=32)
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    binary = v.ExportBVH()
    desc.params.BvhWidth = 4
    v.SetRenderingParams(desc.params)
    v.Update()
    quad = v.ExportBVH()
    bq, bb = _parse_blob(quad), _parse_blob(binary)
    assert bq["hdr"]["flags"] & 2 and not bb["hdr"]["flags"] & 2
    assert bq["hdr"]["n_nodes"] < bb["hdr"]["n_nodes"]
    info, bmin, bmax = bq["info"], bq["bmin"], bq["bmax"]
    counts = []
    def walk(root, off, top):
        stack = [root]
        while stack:
            n = stack.pop()
            x, y, z, w = info[n]
            if x == 0:
                k = z + 1
                assert 2 <= k <= 4
                counts.append(k)
                for c in range(k):
                    ch = off + y + c
                    assert (bmin[ch] >= bmin[n] - 1e-6).all() and (bmax[ch] <= bmax[n] + 1e-6).all()
                    stack.append(ch)
            elif x > 0:
                assert top
                walk(y, y, False)
            else:
                assert not top and 0 < z - y + 1 <= 5
    walk(0, 0, True)
    assert np.mean(counts) > 2.5
    ob, oq = OracleScene(binary), OracleScene(quad)
    org, d = scenes.random_rays(20000, bq["hdr"]["min"], bq["hdr"]["max"], seed=8)
    a = ob.trace(org, d, stats=True)
    b = oq.trace(org, d, stats=True)
    same = (a[0] == b[0]) & (a[1] == b[1])
    assert same.mean() > 0.999                                  # only coplanar near-ties may resolve differently
    assert np.allclose(a[2][same], b[2][same], rtol=1e-6)
    assert b[5]["n_inner"] < 0.9 * a[5]["n_inner"]              # fewer dependent node steps
    assert a[5]["n_boxes"] == 2 * a[5]["n_inner"] and b[5]["n_boxes"] > 2.5 * b[5]["n_inner"]
    s = oq.trace(org, d, any_hit=True)
    assert np.array_equal(s[0] == 0, b[0] >= 0)
    v.Remove()


def test_round1_advice_items(product_lib):
    """Host-side argument checks the round-1 review asked for: non-finite transforms are rejected by
    crt_instance_set_transform too, a hidden instance with a singular transform does not fail the commit,
    attribute arrays of the wrong length never reach crt_mesh_create, BufferDump refuses a wrong `out`."""
    v = V3d_View(host_only=True)
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    m = v.AddMesh(tri, np.array([[0, 1, 2]], np.uint32))
    a = v.Display(m, None)
    b = v.Display(m, scenes.trsf((0, 0, 1)))
    bad = scenes.trsf((0, 0, 0)).copy()
    bad[0, 3] = np.nan
    with pytest.raises(CrtError) as e:
        v.SetLocation(a, bad)
    assert e.value.code == CRT_ERR_INVALID_ARG
    v.Update()                                            # the rejected transform left nothing behind
    v.SetLocation(b, np.zeros((3, 4), np.float32))        # singular ...
    v.SetVisible(b, False)                                # ... but not displayed
    v.Update()
    v.SetVisible(b, True)
    with pytest.raises(CrtError):
        v.Update()                                        # displayed again: the commit refuses it
    v.SetLocation(b, None)
    v.Update()
    with pytest.raises(ValueError):
        v.AddMesh(tri, np.array([[0, 1, 2]], np.uint32), nrm=np.array([[0, 0, 1]], np.float32))
    with pytest.raises(ValueError):
        v.AddMesh(tri, np.array([[0, 1, 2]], np.uint32), uv=np.zeros((2, 2), np.float32))
    v._size = (4, 4)
    with pytest.raises(ValueError):
        v.BufferDump(0, out=np.zeros((2, 2, 3), np.uint8))
    with pytest.raises(ValueError):
        v.BufferDump(0, out=np.zeros((4, 4, 3), np.float32))
    from cadrays_b200.view import Graphic3d_RenderingParams
    assert Graphic3d_RenderingParams().NbRayTracingTiles == 128      # SettingsWidget.cxx:72, same as the C++ mirror
    v.Remove()
