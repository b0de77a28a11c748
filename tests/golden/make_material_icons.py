"""Derives tests/golden/material_icons.json from the reference's own renders.

data/materials/<MaterialName>.png (64x64) are OCCT path-traced renders of data/other/preview.tcl, one per named
material, loaded by the GUI as material buttons (src/Launcher/main.cxx:120-132).  They are the only outputs of the
reference renderer in the repository.  This script stores, per icon, the mean display-space RGB of the ball's centre
disc and of the floor strip -- statistics, not the images.  Run here (the reference tree is mounted read-only):

  python tests/golden/make_material_icons.py
"""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))
from cadrays_b200.imageio import read_png_rgb8  # noqa: E402

SRC = Path("/root/reference/data/materials")


def ball_mask(h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    return ((xx - w / 2 + 0.5) ** 2 + (yy - h * 0.45) ** 2) < (w * 0.18) ** 2


def main():
    out = {"_about": "mean display-space RGB of the ball disc (centre 0.5w,0.45h, radius 0.18w) and the bottom 8 rows of "
                     "data/materials/*.png; made by tests/golden/make_material_icons.py"}
    for f in sorted(SRC.glob("*.png")):
        im = read_png_rgb8(str(f)).astype(np.float64) / 255.0
        h, w, _ = im.shape
        out[f.stem] = {"size": [w, h], "ball": [round(float(v), 4) for v in im[ball_mask(h, w)].mean(0)],
                       "floor": [round(float(v), 4) for v in im[-8:].mean((0, 1))]}
    (Path(__file__).parent / "material_icons.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", len(out) - 1, "icons")


if __name__ == "__main__":
    main()
