"""Interactive cadence of the reference: one sample per pixel per Redraw() (AppViewer.cxx:1047) followed by
the display pass, N frames; prints frames per second like Output_<script>_<N>.txt (main.cxx:224-227)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from cadrays_b200 import scenes
from cadrays_b200.view import V3d_View, Graphic3d_BT_RGB

def run(name, desc, frames=100, spp=1):
    v = V3d_View(0)
    desc.apply(v)
    img = np.empty((desc.height, desc.width, 3), np.uint8)
    for _ in range(5):
        v.Redraw(spp); v.BufferDump(Graphic3d_BT_RGB, img)
    t0 = time.perf_counter()
    for _ in range(frames):
        v.Redraw(spp)
        v.BufferDump(Graphic3d_BT_RGB, img)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    for _ in range(frames):
        v.Redraw(spp)
    dt2 = time.perf_counter() - t1
    print(f"{name:28s} {desc.width}x{desc.height} depth {desc.params.RaytracingDepth} spp/frame {spp}: "
          f"{frames/dt:8.1f} fps with BufferDump, {frames/dt2:8.1f} fps Redraw only ({dt2/frames*1e3:.2f} ms/frame)")
    v.Remove()

if __name__ == "__main__":
    run("cornell C1", scenes.cornell_box(512, 512, depth=5))
    run("materials 1080p", scenes.materials_scene(1920, 1080, depth=12))
    run("assembly C2 1080p", scenes.assembly())
    run("assembly C2 1080p", scenes.assembly(), spp=4)
    run("preview.tcl-like 128x128", scenes.materials_scene(128, 128, depth=10, sphere_res=(64, 32)))
