"""The reference's cadence under ncu: a few Redraw(1) + BufferDump frames of config C2 at 1080p (launch list)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from cadrays_b200 import scenes
from cadrays_b200.view import V3d_View, Graphic3d_BT_RGB

desc = scenes.assembly()
v = V3d_View(0)
desc.apply(v)
img = np.empty((desc.height, desc.width, 3), np.uint8)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    v.Redraw(1)
    v.BufferDump(Graphic3d_BT_RGB, img)
v.Remove()
