"""Generates tests/golden/cornell_golden.npz + golden.json from the CPU oracle (oracle/cadrays_oracle.c)
through the host-only scene builder.  These fixtures pin the ORACLE against accidental change and give the
GPU tests a committed input/output pair; they are not outputs of the reference (OCCT is absent; parity
unpinned).  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from cadrays_b200 import scenes  # noqa: E402
from cadrays_b200.view import V3d_View  # noqa: E402
from oracle.oracle_ffi import OracleScene  # noqa: E402

META = {"width": 48, "height": 40, "depth": 5, "sphere_res": [16, 8], "spp": 3, "n_rays": 4096, "ray_seed": 17}


def main():
    desc = scenes.cornell_box(META["width"], META["height"], depth=META["depth"], sphere_res=tuple(META["sphere_res"]))
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    blob = v.ExportBVH()
    v.Remove()
    o = OracleScene(blob)
    o.configure(desc)
    org, d = scenes.random_rays(META["n_rays"], (0, 0, 0), (1, 1, 1), seed=META["ray_seed"])
    prim, inst, t, u, vv = o.trace(org, d)
    acc = o.render(META["width"], META["height"], META["spp"])
    np.savez_compressed(HERE / "cornell_golden.npz", org=org, dir=d, prim=prim, inst=inst, t=t, u=u, v=vv,
                        accum=acc, ldr=o.display(acc))
    json.dump(META, open(HERE / "golden.json", "w"), indent=1)
    print("wrote", HERE / "cornell_golden.npz")


if __name__ == "__main__":
    main()
