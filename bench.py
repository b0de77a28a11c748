#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its configuration: path-traced Msamples/s
(and Mrays/s) at 1920x1080, depth 8, on the synthetic ~1M-triangle assembly (config C2).

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU oracle on the host cores)

One "step" = one pass of the hot path over one batch: --spp samples per pixel of the whole
frame (generate -> [extend, shade, connect] x depth -> resolve).  With N ranks every rank
holds a scene replica and renders its own disjoint block of --spp sample indices per step
(weak scaling); the float4 accumulation buffers are summed with one NCCL all-reduce per step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "path-traced Msamples/s (1080p, depth 8)"
UNIT = "Msamples/s"


_OUT = sys.stdout


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=16, help="samples per pixel per step and per rank")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--parts", type=int, default=1000)
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--workload", default="assembly", choices=["assembly", "cornell", "materials", "product_shot", "instanced", "instanced_flat"])
    ap.add_argument("--bvh-width", type=int, default=2, choices=[2, 4], help="2 = binary BVH (default), 4 = OCCT's optional QUAD_BVH collapse")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", default="1920x1080x32",
                    help="WxHxSPP sample of the workload for the CPU legs (default: about 5 s per step on 16 cores)")
    return ap.parse_args()


def make_scene(args):
    from cadrays_b200 import scenes
    if args.workload == "assembly":
        return scenes.assembly(n_parts=args.parts, target_tris=args.tris, seed=2, width=args.width, height=args.height, depth=args.depth)
    if args.workload == "cornell":
        return scenes.cornell_box(args.width, args.height, depth=args.depth)
    if args.workload == "materials":
        return scenes.materials_scene(args.width, args.height, depth=args.depth)
    if args.workload == "product_shot":        # C4 geometry and lighting (environment only); pass --width 3840 --height 2160 --depth 12 --spp 4
        d = scenes.product_shot(args.width, args.height, depth=args.depth)
        return d
    if args.workload == "instanced_flat":      # C5, flattened variant: every instance owns its geometry (10.5 M unique triangles)
        return scenes.instanced(n_meshes=1024, width=args.width, height=args.height, depth=args.depth)
    return scenes.instanced(width=args.width, height=args.height, depth=args.depth)


def workload_config(args, desc):
    return {
        "workload": f"C2 synthetic STEP-like assembly: {len(desc.instances)} objects, {desc.n_triangles()} triangles, "
                    f"60% diffuse / 40% glossy, 1 directional light, {args.width}x{args.height}, depth {args.depth}"
                    if args.workload == "assembly" else f"{args.workload} {args.width}x{args.height} depth {args.depth}",
        "spp_per_step_per_gpu": args.spp,
        "bvh_width": args.bvh_width,
        "l2": "per-step working set (about 5 GB of path state + 0.12 GB of scene) exceeds the 126 MB L2; no explicit flush",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu), "-f", self.path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower() == "active":
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def algorithmic_bytes(st: dict):
    """an inner visit reads 16 B of node info + 24 B per child box tested (= 64 B for a binary node),
    16 B per leaf visit, 52 B per triangle test, 64 B per level switch."""
    near = 16 * st["n_inner"] + 24 * st["n_boxes"] + 16 * st["n_leaf"] + 52 * st["n_tri"] + 64 * st["n_switch"]
    anyh = 16 * st["n_inner_any"] + 24 * st["n_boxes_any"] + 16 * st["n_leaf_any"] + 52 * st["n_tri_any"] + 64 * st["n_switch_any"]
    shade = (36 + 128) * st["shaded_hits"] + 32 * st["samples"]
    return near, anyh, shade


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per k_extend launch from the committed ncu --set full capture, or None."""
    p = REPO / "profiles" / "extend_traffic.json"
    if p.exists():
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def cpu_leg(args, desc, blob, nthreads, steps):
    """Times the CPU oracle (restated reference algorithm) on a bounded sample of the same workload."""
    from oracle.oracle_ffi import OracleScene
    w, h, spp = (int(v) for v in args.cpu_sample.lower().split("x"))
    orc = OracleScene(blob)
    orc.configure(desc)   # same scene, camera and parameters as the GPU arm
    import numpy as np
    times = []
    for s in range(max(1, steps)):
        acc = np.zeros((h, w, 4), dtype=np.float32)
        t0 = time.perf_counter()
        orc.render(w, h, spp, first_sample=s * spp, accum=acc, nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    orc.close()
    dt = sum(times) / len(times)
    return (w * h * spp) / dt / 1e6, dt, f"{w}x{h} x {spp} spp of the same view and scene per step, {len(times)} steps"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference renderer is OCCT's
    GLSL under Mesa llvmpipe, which cannot be built or run here (OCCT, Tcl, GL stack absent); the CPU oracle
    port of the same algorithm is timed instead, on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cadrays_b200.view import V3d_View
    desc = make_scene(args)
    host = V3d_View(host_only=True)
    desc.apply(host, with_target=False)
    blob = host.ExportBVH()
    host.Remove()
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_leg(args, desc, blob, cores, 1)
    v, dt, sample = cpu_leg(args, desc, blob, cores, min(args.steps, 5))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": min(args.steps, 5),
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, desc),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle port (OpenMP over pixel rows); the OCCT/llvmpipe reference binary is not runnable here",
    }
    _OUT.write(json.dumps(line) + "\n"); _OUT.flush()


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cadrays_b200 import distributed as D
    from cadrays_b200.view import Graphic3d_BT_RGB, V3d_View

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcadrays_b200 has no CPU fallback")
    # NCCL logs (version banner included) go to stdout by default: send them to stderr so that stdout
    # carries exactly one JSON line whatever NCCL_DEBUG level the caller chose
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, local, world = D.init_from_env("nccl")
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    desc = make_scene(args)
    t_build0 = time.perf_counter()
    view = V3d_View(local)
    desc.apply(view)
    t_build = time.perf_counter() - t_build0
    W, H, B, depth = args.width, args.height, args.spp, args.depth
    p = desc.params
    p.SamplesPerBatch = B
    p.BvhWidth = args.bvh_width
    view.SetRenderingParams(p)
    view.Update()

    stream = torch.cuda.ExternalStream(view.Stream(), device=torch.device("cuda", local))
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=f"cuda:{local}")
    reduced = torch.zeros_like(accum)
    torch.cuda.synchronize()
    view.BindAccum(accum.data_ptr(), accum.numel() * 4)

    # the all-reduce of step s runs on its own stream while the context stream already traces step s + 1:
    # snapshot the sums (device copy, context stream), hand the snapshot to the communication stream, and let the
    # next snapshot wait until the previous all-reduce has consumed the buffer
    comm = torch.cuda.Stream(device=torch.device("cuda", local)) if world > 1 else None
    reduce_done = torch.cuda.Event() if world > 1 else None
    snapshot_ready = torch.cuda.Event() if world > 1 else None

    def one_step(step_index):
        view.SetNextSample(D.step_sample_start(step_index, rank, world, B))
        view.RedrawAsync(B)
        if world > 1:
            stream.wait_event(reduce_done)           # no-op before the first record
            reduced.copy_(accum, non_blocking=True)
            snapshot_ready.record(stream)
            comm.wait_event(snapshot_ready)
            with torch.cuda.stream(comm):
                dist.all_reduce(reduced)
                reduce_done.record(comm)

    def join_comm():
        if world > 1:
            stream.wait_event(reduce_done)           # the timed region ends after the last all-reduce

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for s in range(args.warmup):
            one_step(s)
        barrier()
        # ---- timed region: device time on the launching stream, per-kernel spans inside
        view.ResetStats()
        view.EnableTiming(True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for s in range(args.steps):
            one_step(args.warmup + s)
        join_comm()
        ev1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = ev0.elapsed_time(ev1)
        timing = view.Timing()
        view.EnableTiming(False)
        t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

        # ---- the same steps again with the instrumented kernels: work counters (untimed)
        view.EnableStats(True)
        view.ResetStats()
        for s in range(args.steps):
            view.SetNextSample(D.step_sample_start(args.warmup + s, rank, world, B))
            view.RedrawAsync(B)
        view.Sync()
        stats = view.Stats()
        view.EnableStats(False)

    samples_per_step = W * H * B * world
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6
    rays = stats["rays_nearest"] + stats["rays_any"]
    mrays = rays * world / (ms_total * 1e-3) / 1e6

    # ---- e2e: the public call a user makes, host buffers, copies inside the timed region
    view.BindAccum(None)
    ldr = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy()    # pinned host frame buffer
    cam = desc.camera
    for s in range(2):
        view.SetCamera(cam); view.Redraw(B); view.BufferDump(Graphic3d_BT_RGB, ldr)
    e2e_steps = max(3, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        view.SetCamera(cam)                  # host -> device: camera / params (kernel arguments) + frame seeds
        view.Redraw(B)                       # V3d_View::Redraw
        view.BufferDump(Graphic3d_BT_RGB, ldr)   # device -> host: tone-mapped RGB8 frame
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = W * H * B * world * e2e_steps / float(te.item()) / 1e6

    if rank == 0:
        near_b, any_b, shade_b = algorithmic_bytes(stats)
        ext_ms, ext_n = timing["extend"]
        con_ms, con_n = timing["connect"]
        peak, peak_src = measured_peak()
        # traversal launches = extend family (closest hit, and the fused closest+any-hit launches) + connect
        # family (remaining any-hit launches); their algorithmic bytes are the closest-hit + any-hit counters
        trav_ms, trav_n = ext_ms + con_ms, ext_n + con_n
        trav_b = near_b + any_b
        achieved = (trav_b / (trav_ms * 1e-3)) / 1e9 if trav_ms > 0 else 0.0
        nr, na = max(stats["rays_nearest"], 1), max(stats["rays_any"], 1)
        roofline = {
            "bound": "hbm", "kernel": "traversal launches: k_extend + k_trace_dual + k_connect (SceneNearestHit / SceneAnyHit)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": ncu_traffic(),
            "algorithmic_bytes_per_launch": trav_b / max(trav_n, 1), "launch_ms": trav_ms / max(trav_n, 1), "launches": trav_n,
            "algorithmic_bytes_per_step": trav_b / args.steps, "traversal_ms_per_step": trav_ms / args.steps,
            "per_ray_nearest": {"n_inner": stats["n_inner"] / nr, "n_boxes": stats["n_boxes"] / nr, "n_leaf": stats["n_leaf"] / nr, "n_tri": stats["n_tri"] / nr, "n_switch": stats["n_switch"] / nr},
            "per_ray_any": {"n_inner": stats["n_inner_any"] / na, "n_leaf": stats["n_leaf_any"] / na, "n_tri": stats["n_tri_any"] / na, "n_switch": stats["n_switch_any"] / na},
            "note": "algorithmic bytes use the reference's record sizes (SURVEY 8(d): 64 B inner visit, 16 B leaf, 52 B triangle, 64 B switch); "
                    "the scene is largely L1/L2-resident, so achieved may exceed the HBM copy peak -- see profiles/ for what binds",
        }
        kernel_ms = {k: v[0] / args.steps for k, v in timing.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, desc),
            "mrays_per_s": mrays, "rays_per_sample": rays / max(stats["samples"], 1),
            "roofline": roofline, "kernel_ms_per_step": kernel_ms,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": W * H * 3,
                    "steps": e2e_steps, "call": "SetCamera + Redraw(spp) + BufferDump(RGB8) per step, wall clock"},
            "gpu_launches": int(sum(v[1] for k, v in timing.items() if k != "render")),
            "scene_commit_s": t_build,
        }
        if not args.no_cpu_baseline and world == 1:
            blob = view.ExportBVH()
            v, dtc, sample = cpu_leg(args, desc, blob, os.cpu_count() or 1, 3)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
        _OUT.write(json.dumps(line) + "\n"); _OUT.flush()
    view.Remove()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout must carry exactly ONE JSON line.  Native libraries (NCCL's version banner, for one) print to
    # file descriptor 1 directly, so route fd 1 to stderr for the whole run and keep the real stdout for us.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
