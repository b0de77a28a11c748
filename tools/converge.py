"""Convergence runs of the BASELINE configs that are quoted on several GPUs:
  C3  Materials.tcl scene, depth 12, 1080p, samples partitioned over the ranks
  C4  environment-lit product shot, 3840x2160, 4096 spp at 1/2/4/8 GPUs

One process per GPU (plain `python` for one GPU, torchrun for more); rank r renders the r-th block of the
sample range, the float sum buffers are added with one NCCL all-reduce per reporting interval, rank 0 prints
one JSON line.  With --compare the N-GPU sum is checked against the 1-GPU image written earlier on the same box
(same sample set, only the order of the float additions differs).

  python tools/converge.py --config c4 --spp 4096 --save /tmp/c4_n1.npy
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      tools/converge.py --config c4 --spp 4096 --compare /tmp/c4_n1.npy
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=["c2", "c3", "c4"])
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--interval", type=int, default=0, help="samples per rank between all-reduces (0 = one at the end)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--save", default="")
    ap.add_argument("--compare", default="")
    args = ap.parse_args()
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                      # NCCL banners etc. go to stderr; stdout carries the JSON line only
    import torch
    import torch.distributed as dist
    from cadrays_b200 import distributed as D
    from cadrays_b200 import scenes
    from cadrays_b200.view import V3d_View

    rank, local, world = D.init_from_env("nccl")
    torch.cuda.set_device(local)
    if args.config == "c4":
        desc = scenes.product_shot()
    elif args.config == "c3":
        desc = scenes.materials_scene(1920, 1080, depth=12)
    else:
        desc = scenes.assembly()
    W, H = desc.width, desc.height
    view = V3d_View(local)
    desc.apply(view)
    p = desc.params
    p.SamplesPerBatch = args.batch or max(1, min(16, (32 << 20) // (W * H)))
    view.SetRenderingParams(p)
    view.Update()
    stream = torch.cuda.ExternalStream(view.Stream(), device=torch.device("cuda", local))
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=f"cuda:{local}")
    reduced = torch.zeros_like(accum)
    torch.cuda.synchronize()
    view.BindAccum(accum.data_ptr(), accum.numel() * 4)
    first, count = D.sample_range(rank, world, args.spp)
    interval = args.interval or count

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        view.ResetAccumulation(1 << 24)           # warm-up on a sample range that is not part of the run
        view.RedrawAsync(p.SamplesPerBatch)
        if world > 1:
            reduced.copy_(accum); dist.all_reduce(reduced)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        view.ResetAccumulation(first)
        done = 0
        n_reduce = 0
        while done < count:
            n = min(interval, count - done)
            view.RedrawAsync(n)
            done += n
            reduced.copy_(accum, non_blocking=True)
            if world > 1:
                dist.all_reduce(reduced)
            n_reduce += 1
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    line = None
    if rank == 0:
        total = reduced.cpu().numpy()
        assert float(total[..., 3].min()) == float(total[..., 3].max()) == float(args.spp), "sample counts do not add up"
        mean = total[..., :3] / total[..., 3:4]
        line = {"config": args.config, "size": [W, H], "depth": int(p.RaytracingDepth), "spp": args.spp, "n_gpus": world,
                "seconds": ms * 1e-3, "msamples_per_s": W * H * args.spp / (ms * 1e-3) / 1e6,
                "allreduces": n_reduce if world > 1 else 0, "allreduce_bytes": int(accum.numel() * 4),
                "samples_per_batch": int(p.SamplesPerBatch), "mean_radiance": float(mean.mean()),
                "finite": bool(np.isfinite(mean).all())}
        if args.save:
            np.save(args.save, total)
        if args.compare and os.path.exists(args.compare):
            ref = np.load(args.compare)
            rm = ref[..., :3] / ref[..., 3:4]
            rel = np.abs(mean - rm) / np.maximum(np.abs(rm), 1e-3)
            d8 = lambda m: (np.sqrt(np.clip(m, 0, 1)) * 255.0 + 0.5).astype(np.uint8)
            line["vs_1gpu"] = {"max_rel_diff": float(rel.max()), "rmse": float(np.sqrt(np.mean((mean - rm) ** 2))),
                               "rgb8_values_differing": int(np.sum(d8(mean) != d8(rm))), "rgb8_values": int(mean.size)}
    view.BindAccum(None)
    view.Remove()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        real_out.write(json.dumps(line) + "\n")
        real_out.flush()


if __name__ == "__main__":
    main()
