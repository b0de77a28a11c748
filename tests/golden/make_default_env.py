"""Decodes the reference's default environment map (data/maps/default.jpg, the map CADRays loads at
AppGui.cxx:963 and BASELINE config C4 is lit by) into a lossless fixture that travels to the GPU box,
where /root/reference does not exist.  Run in the build container:

    python tests/golden/make_default_env.py

The JPEG is greyscale content stored as RGB (r = g = b in every pixel; checked below), so the fixture is
an 8-bit single-channel PNG of the decoded pixels; scenes.default_env() expands it back to (h, w, 3) uint8.
"""
import sys
from pathlib import Path

import numpy as np
from PIL import Image

SRC = Path("/root/reference/data/maps/default.jpg")
DST = Path(__file__).resolve().parent / "default_env_2048x1024.png"

if __name__ == "__main__":
    a = np.asarray(Image.open(SRC).convert("RGB"))
    assert a.shape == (1024, 2048, 3), a.shape
    assert np.array_equal(a[..., 0], a[..., 1]) and np.array_equal(a[..., 0], a[..., 2]), "not greyscale any more: store RGB"
    Image.fromarray(a[..., 0], mode="L").save(DST, "PNG", optimize=True)
    back = np.asarray(Image.open(DST))
    assert np.array_equal(back, a[..., 0])
    print("wrote", DST, DST.stat().st_size, "bytes; mean", float(a.mean()))
    sys.exit(0)
