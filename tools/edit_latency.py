"""Edit-to-first-sample latency of an instance-only edit at config C2 (1 M triangles): SetLocation (the gizmo drag of
ImRaytraceControls.cxx:88) / material assignment (MaterialEditor.cxx:522-523), then Update + Redraw(1) + BufferDump.
Prints one JSON object.  Usage: python tools/edit_latency.py [--devices 0,1]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="")
    ap.add_argument("--edits", type=int, default=20)
    args = ap.parse_args()
    from cadrays_b200 import scenes
    from cadrays_b200.view import Graphic3d_BT_RGB, V3d_View
    desc = scenes.assembly()
    devices = [int(d) for d in args.devices.split(",")] if args.devices else None
    view = V3d_View(0) if devices is None else V3d_View(devices=devices)
    t0 = time.perf_counter()
    desc.apply(view)
    t_full = time.perf_counter() - t0
    view.Redraw(1)
    ldr = np.empty((desc.height, desc.width, 3), np.uint8)
    view.BufferDump(Graphic3d_BT_RGB, ldr)
    out = {"workload": "C2 assembly, %d triangles, %d objects" % (desc.n_triangles(), len(desc.instances)),
           "members": len(devices) if devices else 1, "first_commit_s": t_full}
    ms_commit, ms_total = [], []
    before = view.CommitStats()
    for k in range(args.edits):
        inst = 5 + 7 * k
        xf = np.array(desc.instances[inst][1], dtype=np.float32).reshape(3, 4).copy()
        xf[:, 3] += (0.05 * (k % 3), 0.03, 0.02 * (k % 5))
        t0 = time.perf_counter()
        view.SetLocation(inst, xf.reshape(12))
        if k % 2:
            view.SetMaterialIndex(inst + 1, (k * 13) % max(1, len(desc.materials)))
        view.Update()
        t1 = time.perf_counter()
        view.Redraw(1)
        view.BufferDump(Graphic3d_BT_RGB, ldr)
        t2 = time.perf_counter()
        ms_commit.append((t1 - t0) * 1e3)
        ms_total.append((t2 - t0) * 1e3)
    out["top_level_patches"] = view.CommitStats() - before
    out["edits"] = args.edits
    out["commit_ms_median"] = float(np.median(ms_commit))
    out["edit_to_first_frame_ms_median"] = float(np.median(ms_total))
    out["edit_to_first_frame_ms_max"] = float(np.max(ms_total))
    view.Remove()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
