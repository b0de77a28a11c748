"""CPU estimate of what a builder variant does to the traversal work of a config: builds the scene with the host code
(no GPU), walks it with the oracle (a w x h x spp path-traced frame with work counters) and prints the counters per
ray plus a cost in 'node-visit units' calibrated on the ncu source profile of k_trace_dual (one triangle test = 1.8
node visits, one instance switch = 1.25).  usage: python tests/analysis/tree_quality.py [workload] KNOB=VALUE ...   (one
setting per process: the builder reads its knobs once)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=", 1); os.environ[k] = v
from cadrays_b200 import scenes
from cadrays_b200.view import V3d_View
from oracle.oracle_ffi import OracleScene
which = next((a for a in sys.argv[1:] if "=" not in a), "assembly")
desc = {"assembly": scenes.assembly, "instanced": scenes.instanced, "cornell": scenes.cornell_box, "materials": scenes.materials_scene}[which]()
v = V3d_View(host_only=True)
t0 = time.perf_counter(); desc.apply(v, with_target=False); blob = v.ExportBVH(); tb = time.perf_counter() - t0
orc = OracleScene(blob); orc.configure(desc)
w, h = 480, 270
orc.set_camera_aspect(w / h) if hasattr(orc, "set_camera_aspect") else None
_, st = orc.render(w, h, 1, stats=True)
rays = st["rays_nearest"] + st["rays_any"]
ni = (st["n_inner"] + st["n_inner_any"]) / rays; nt = (st["n_tri"] + st["n_tri_any"]) / rays
nl = (st["n_leaf"] + st["n_leaf_any"]) / rays; ns = (st["n_switch"] + st["n_switch_any"]) / rays
print(f"{which} {[a for a in sys.argv[1:] if '=' in a]}: blob {len(blob)/1e6:.0f} MB build {tb:.1f}s | per ray: inner {ni:.2f} leaf {nl:.2f} tri {nt:.2f} switch {ns:.2f} | cost {ni + 1.8*nt + 1.25*ns:.2f}")
