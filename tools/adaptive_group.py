"""Adaptive screen sampling through a crt_group: samples per second of config C2 at 1080p with 1, 2, ... GPUs driven from
one host process (V3d_View(devices=[...]), AdaptiveScreenSampling on), plain sampling beside it.
usage: python tools/adaptive_group.py [spp] -> one JSON line"""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from cadrays_b200 import scenes
from cadrays_b200.view import V3d_View, Graphic3d_BT_RGB

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {"workload": "C2 assembly 1920x1080 depth 8", "spp": spp, "runs": []}
n_gpus = torch.cuda.device_count()
for n in [k for k in (1, 2, 4, 8) if k <= n_gpus]:
    for adaptive in (False, True):
        desc = scenes.assembly()
        desc.params.AdaptiveScreenSampling = adaptive
        desc.params.NbRayTracingTiles = 0
        v = V3d_View(devices=list(range(n)))
        desc.apply(v)
        img = torch.empty((desc.height, desc.width, 3), dtype=torch.uint8, pin_memory=True).numpy()
        v.Redraw(spp); v.BufferDump(Graphic3d_BT_RGB, img)      # warm-up with the timed call's own sizes (path state, seed tables)
        v.SetCamera(desc.camera)                               # restarts the accumulation on every member
        v.Redraw(1)
        v.SetCamera(desc.camera)
        t0 = time.perf_counter()
        v.Redraw(spp); v.BufferDump(Graphic3d_BT_RGB, img)
        dt = time.perf_counter() - t0
        counts = v.SamplingTiles()[0] if adaptive else None
        out["runs"].append({"gpus": n, "adaptive": adaptive, "seconds": dt, "msamples_per_s": desc.width * desc.height * spp / dt / 1e6,
                            "tile_samples_min_max": [int(counts.min()), int(counts.max())] if adaptive else None})
        v.Remove()
print(json.dumps(out))
