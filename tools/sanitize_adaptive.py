"""Small adaptive-sampling + pipelined render for compute-sanitizer runs (memcheck / racecheck / initcheck)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cadrays_b200 import scenes
from cadrays_b200.view import V3d_View

desc = scenes.cornell_box(101, 67, depth=4, sphere_res=(16, 8))
p = desc.params
p.AdaptiveScreenSampling, p.NbRayTracingTiles, p.SamplesPerBatch = True, 5, 2
v = V3d_View(0)
desc.apply(v)
v.Redraw(7)
print("adaptive tiles", v.SamplingTiles()[0].reshape(-1).tolist())
p.AdaptiveScreenSampling = False
p.SamplesPerBatch = 3
v.SetRenderingParams(p)
v.Redraw(6)
print("plain ok", v.BufferDump().mean())
v.Remove()
