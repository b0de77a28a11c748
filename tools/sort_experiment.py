"""Upper bound of what a GLOBAL ray sort could buy the depth-1 traversal.

Takes the real bounce-1 continuation rays and bounce-0 shadow rays of config C2 out of the wavefront
(crt_wavefront_rays), and times crt_trace_device on the same rays in three orders:
  wavefront  the order the path tracer traces them in today (slot order = 8x4 pixel tiles, compacted)
  sorted     globally sorted by (direction octant, 30-bit Morton code of the origin) -- what a
             cub::DeviceRadixSort of the queue would produce, WITHOUT its cost and without the scattered
             path-state reads it would cause
  shuffled   a random permutation (the incoherent extreme)
Prints one JSON object.  Usage: python tools/sort_experiment.py [--workload assembly|instanced_flat] [--spp 4]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def morton30(p, lo, hi):
    q = np.clip((p - lo) / np.maximum(hi - lo, 1e-20), 0, 1 - 1e-7)
    q = (q * 1024).astype(np.uint64)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="assembly")
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from cadrays_b200 import scenes
    from cadrays_b200.view import V3d_View

    if args.workload == "assembly":
        desc = scenes.assembly()
    else:
        desc = scenes.instanced(n_meshes=1024)
    desc.params.RaytracingDepth = 2
    desc.params.SamplesPerBatch = args.spp
    view = V3d_View(0)
    desc.apply(view)
    view.Redraw(args.spp)
    out = {"workload": args.workload, "spp": args.spp}
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(view.Stream(), device=dev)
    for kind, shadow, depth in (("continuation_d1", False, 1), ("shadow_d0", True, 0)):
        org, d, tmax = view.WavefrontRays(depth, shadow=shadow)
        n = org.shape[0]
        lo, hi = org.min(0), org.max(0)
        octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2)
        key = (octant << 30) | morton30(org.astype(np.float64), lo, hi)
        orders = {"wavefront": np.arange(n), "sorted": np.argsort(key, kind="stable"),
                  "shuffled": np.random.default_rng(1).permutation(n)}
        res = {"rays": int(n)}
        for name, perm in orders.items():
            o4 = np.zeros((n, 4), np.float32); o4[:, :3] = org[perm]
            d4 = np.zeros((n, 4), np.float32); d4[:, :3] = d[perm]; d4[:, 3] = tmax[perm]
            to, td = torch.from_numpy(o4).to(dev), torch.from_numpy(d4).to(dev)
            hit = torch.empty((n, 4), dtype=torch.float32, device=dev)
            inst = torch.empty(n, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            ms = []
            with torch.cuda.stream(stream):
                for r in range(args.reps + 1):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    view.TraceDevice(to.data_ptr(), td.data_ptr(), n, hit.data_ptr(), inst.data_ptr(), any_hit=shadow)
                    e1.record(stream)
                    view.Sync()
                    if r:
                        ms.append(e0.elapsed_time(e1))
            res[name] = {"ms": float(np.median(ms)), "mrays_per_s": n / float(np.median(ms)) / 1e3}
            del to, td, hit, inst
        res["sorted_over_wavefront"] = res["wavefront"]["ms"] / res["sorted"]["ms"]
        out[kind] = res
    view.Remove()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
