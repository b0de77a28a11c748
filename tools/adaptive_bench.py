"""Adaptive screen sampling (crt_params.adaptive_sampling) measured on the GPU: throughput against the plain
mode at the same path budget, and the display-space error both reach against a long plain render.

  python tools/adaptive_bench.py [--scene cornell|assembly|materials] [--budget 64] [--ref 2048]

Prints one JSON line.  Analysis tooling: not part of bench.py's contract.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def display(hdr):
    return np.sqrt(np.clip(hdr, 0.0, 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--budget", type=int, default=64, help="samples per pixel (plain) / tile samples per tile (adaptive)")
    ap.add_argument("--ref", type=int, default=2048, help="samples per pixel of the plain reference image")
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    import torch
    from cadrays_b200 import scenes
    from cadrays_b200.view import Graphic3d_BT_RGB_RayTraceHdrLeft, V3d_View
    if args.scene == "cornell":
        desc = scenes.cornell_box(512, 512, depth=5)
    elif args.scene == "materials":
        desc = scenes.materials_scene(1920, 1080, depth=12)
    else:
        desc = scenes.assembly()
    w, h = desc.width, desc.height
    batch = args.batch or max(1, min(16, (32 << 20) // (w * h)))
    view = V3d_View(0)
    p = desc.params
    p.SamplesPerBatch = batch
    desc.apply(view)
    view.SetRenderingParams(p)
    view.Update()

    def timed(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        view.Redraw(n)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    # reference: a disjoint sample range, so it is independent of both candidates
    view.ResetAccumulation(1 << 20)
    timed(args.ref)
    ref = display(view.BufferDump(Graphic3d_BT_RGB_RayTraceHdrLeft).copy())

    out = {"scene": args.scene, "size": [w, h], "budget": args.budget, "ref_spp": args.ref, "batch": batch}
    for mode in ("plain", "adaptive"):
        p.AdaptiveScreenSampling = mode == "adaptive"
        p.NbRayTracingTiles = 0
        view.SetRenderingParams(p)
        view.ResetAccumulation(0)
        timed(batch)                      # warm-up (allocations), then restart
        view.ResetAccumulation(0)
        view.EnableStats(False)
        dt = timed(args.budget)
        img = display(view.BufferDump(Graphic3d_BT_RGB_RayTraceHdrLeft).copy())
        # the same render again with the instrumented kernels: rays traced and per-family device time
        view.ResetAccumulation(0)
        view.EnableStats(True); view.ResetStats()
        view.Redraw(args.budget)
        st = view.Stats()
        view.EnableStats(False)
        view.ResetAccumulation(0)
        view.ResetStats()
        view.EnableTiming(True)
        view.Redraw(args.budget)
        fam = view.Timing()
        view.EnableTiming(False)
        err = img - ref
        rec = {"seconds": dt, "rmse_display": float(np.sqrt(np.mean(err ** 2))),
               "p99_abs_display": float(np.quantile(np.abs(err), 0.99)),
               "msamples_per_s": w * h * args.budget / dt / 1e6,
               "mrays_per_s": (st["rays_nearest"] + st["rays_any"]) / dt / 1e6,
               "rays_per_sample": (st["rays_nearest"] + st["rays_any"]) / max(1, st["samples"]),
               "samples": st["samples"], "family_ms": {k: round(v[0], 3) for k, v in fam.items()}}
        if mode == "adaptive":
            counts, errs = view.SamplingTiles()
            rec["tile_samples"] = {"min": int(counts.min()), "median": float(np.median(counts)), "max": int(counts.max()),
                                   "sum": int(counts.sum()), "tiles": int(counts.size)}
        out[mode] = rec
        tile_rmse = []
        e2 = (err ** 2).mean(axis=2)
        for ty in range((h + 31) // 32):
            for tx in range((w + 31) // 32):
                tile_rmse.append(float(np.sqrt(e2[ty * 32:(ty + 1) * 32, tx * 32:(tx + 1) * 32].mean())))
        rec["worst_tile_rmse"] = max(tile_rmse)
        rec["p95_tile_rmse"] = float(np.quantile(tile_rmse, 0.95))
    view.Remove()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
