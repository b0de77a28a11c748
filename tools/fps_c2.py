"""Frames per second of config C2 at the reference's cadence (Redraw(1) + BufferDump), one line per call."""
import sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from cadrays_b200 import scenes
from cadrays_b200.view import V3d_View, Graphic3d_BT_RGB
import torch
which = sys.argv[1] if len(sys.argv) > 1 else "assembly"
desc = scenes.assembly() if which == "assembly" else scenes.materials_scene(1920, 1080, depth=12) if which == "materials" else scenes.cornell_box(512, 512, depth=5)
v = V3d_View(0)
desc.apply(v)
img = torch.empty((desc.height, desc.width, 3), dtype=torch.uint8, pin_memory=True).numpy()
for _ in range(10):
    v.Redraw(1); v.BufferDump(Graphic3d_BT_RGB, img)
frames = 200
t0 = time.perf_counter()
for _ in range(frames):
    v.Redraw(1); v.BufferDump(Graphic3d_BT_RGB, img)
dt = time.perf_counter() - t0
t1 = time.perf_counter()
for _ in range(frames):
    v.Redraw(1)
dt2 = time.perf_counter() - t1
knobs = {k: os.environ[k] for k in os.environ if k.startswith("CRT_")}
print(f"{which} {knobs}: {frames/dt:.1f} fps with BufferDump ({dt/frames*1e3:.3f} ms), Redraw only {dt2/frames*1e3:.3f} ms")
v.Remove()
