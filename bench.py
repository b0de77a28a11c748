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
                    help="WxHxSPP sample of the workload for the cpu_baseline leg (default: about 5 s per step on 16 cores)")
    ap.add_argument("--ref-budget-s", type=float, default=75.0,
                    help="--impl reference: CPU seconds the K timed + W warm-up steps may take together; the samples per "
                         "pixel of one step are sized from a calibration pass so that the run fits")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra blocks of the N=1 line (memory probe, e2e_1spp, c5_flattened)")
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


def algorithmic_instructions(st: dict):
    """SURVEY 8(d), issue bound: thread-level SASS instructions a ray needs, instr(ray) = 45 n_inner + 55 n_tri +
    60 n_switch + 30, summed over the closest-hit and any-hit rays of the counters."""
    rays = st["rays_nearest"] + st["rays_any"]
    return (45 * (st["n_inner"] + st["n_inner_any"]) + 55 * (st["n_tri"] + st["n_tri_any"])
            + 60 * (st["n_switch"] + st["n_switch_any"]) + 30 * rays)


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _profile_json(name):
    p = REPO / "profiles" / name
    if p.exists():
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def ncu_traffic(workload):
    """dram bytes per traversal launch from the committed ncu capture of this workload (with the commit and the
    command that produced it), or None."""
    d = _profile_json("extend_traffic.json")
    if not d:
        return None, None
    entry = d.get(workload) if isinstance(d.get(workload), dict) else (d if workload == "assembly" and "dram_bytes_per_launch" in d else None)
    if not entry:
        return None, None
    prov = {k: entry.get(k) for k in ("source", "commit", "command", "round", "launches") if k in entry}
    return entry.get("dram_bytes_per_launch"), prov


def build_roofline(workload, stats, timing, steps, scene_bytes, clocks, probe_dev=None, serial_pass=None):
    """The traversal launches (SceneNearestHit / SceneAnyHit) against every bound SURVEY 8(d) names: HBM by
    algorithmic bytes (the contract's achieved / peak / frac), real DRAM traffic from ncu, measured L2 bandwidth,
    and the issue bound; `bound` names the unit that the counters show closest to its limit."""
    near_b, any_b, _ = algorithmic_bytes(stats)
    ext_ms, ext_n = timing["extend"]
    con_ms, con_n = timing["connect"]
    peak, peak_src = measured_peak()
    # traversal launches = extend family (closest hit, and the fused closest + any-hit launches) + connect family
    # (remaining any-hit launches); their algorithmic bytes are the closest-hit + any-hit counters
    trav_ms, trav_n = ext_ms + con_ms, ext_n + con_n
    trav_b = near_b + any_b
    secs = max(trav_ms, 1e-9) * 1e-3
    achieved = trav_b / secs / 1e9
    nr, na = max(stats["rays_nearest"], 1), max(stats["rays_any"], 1)
    rays = stats["rays_nearest"] + stats["rays_any"]
    traffic, traffic_prov = ncu_traffic(workload)
    hbm = {"achieved": achieved, "peak": peak, "frac": achieved / peak, "unit": "GB/s", "peak_source": peak_src,
           "basis": "algorithmic bytes (reference record sizes) / traversal time"}
    if traffic:
        hbm["dram_traffic_gbs"] = traffic / (trav_ms / max(trav_n, 1) * 1e-3) / 1e9
        hbm["dram_frac"] = hbm["dram_traffic_gbs"] / peak
        hbm["traffic_over_algorithmic"] = traffic * trav_n / max(trav_b, 1)
    # ---- L2: measured on this device, now (cadrays_b200/probe.py)
    l2 = None
    if probe_dev is not None:
        try:
            from cadrays_b200 import probe
            info = probe.device_info(probe_dev)
            ws = int(min(max(scene_bytes, 8 << 20), 8 << 30))
            seq = probe.bandwidth(probe_dev, ws, "sequential")
            rec = probe.bandwidth(probe_dev, ws, "records")
            dep = probe.bandwidth(probe_dev, ws, "records_dependent")
            resident = ws <= info["l2_bytes"]
            l2 = {"capacity_mb": info["l2_bytes"] / 2**20, "working_set_mb": ws / 2**20,
                  "working_set": "nodes + triangle vertices + instance records of this scene" + ("" if resident else " (larger than L2: the figures below are HBM figures)"),
                  "l2_resident": resident,
                  "peak_sequential": seq, "peak_records64": rec, "peak_records64_dependent": dep, "unit": "GB/s",
                  "achieved": achieved, "frac": achieved / rec if rec > 0 else None,
                  "frac_of": "scattered 64-byte record reads over a working set of this size, measured on this device in this run",
                  "note": "records64 = every lane reads its own 64-byte record at an independent random index (two 256-bit loads), the shape "
                          "of a node fetch; _dependent = next index derived from the loaded record (the latency-bound form of the same walk)"}
        except Exception as e:  # the probe must never take the bench down
            l2 = {"error": str(e)}
    # ---- issue: thread instructions the rays need / (32 lanes x issue slots per second)
    cal = _profile_json("issue_calibration.json") or {}
    ent = cal.get(workload, {})
    k = float(ent.get("thread_inst_per_formula_unit", 1.0))
    sm = 148
    if l2 and "capacity_mb" in l2:
        try:
            from cadrays_b200 import probe
            sm = probe.device_info(probe_dev)["sm_count"]
        except Exception:
            pass
    clk_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    slots = sm * 4 * clk_mhz * 1e6                                   # warp instructions per second the schedulers can issue
    t_inst = algorithmic_instructions(stats) * k                     # thread-level instructions of the timed steps
    issue = {"formula": "45 n_inner + 55 n_tri + 60 n_switch + 30 thread instructions per ray (SURVEY 8(d))",
             "calibration": ent or None,
             "thread_inst_per_ray": t_inst / max(rays, 1),
             "achieved": t_inst / 32.0 / secs / 1e9, "peak": slots / 1e9, "unit": "G warp-instructions/s",
             "frac": t_inst / 32.0 / secs / slots,
             "frac_of": f"{sm} SMs x 4 schedulers x {clk_mhz:.0f} MHz (the clock sampled during the timed region); lane-weighted: "
                        "the issue slots the rays would need if every warp instruction ran with 32 useful lanes",
             "issue_active_pct": ent.get("issue_active_pct"), "lanes_per_inst": ent.get("lanes_per_inst"),
             "l1_data_pipe_pct": ent.get("l1_data_pipe_pct")}
    # ---- which unit binds: the one the ncu counters of the committed capture show nearest its limit
    # hardware utilisation of each unit during the traversal launches (ncu, time-weighted over one step; the L2 figure
    # is lts__throughput -- roofline.l2.frac is a different thing: algorithmic bytes over the measured record bandwidth)
    util = {"issue": ent.get("issue_active_pct"), "l1": ent.get("l1_data_pipe_pct"), "l2": ent.get("l2_pct"),
            "hbm": (100.0 * hbm["dram_frac"]) if "dram_frac" in hbm else None}
    known = {u: v for u, v in util.items() if isinstance(v, (int, float))}
    bound = max(known, key=known.get) if known else "unknown"
    return {
        "bound": bound, "utilisation_pct": util,
        "bound_note": "`bound` is the unit the ncu counters of the committed launch list (profiles/issue_calibration.json) show nearest "
                      "its limit during the traversal launches: the L1 data pipe (one wavefront per lane and node / triangle / stack "
                      "access), with issue slots half busy at 17 of 32 lanes; hbm / l2 / issue give every fraction SURVEY 8(d) asks for",
        "kernel": "traversal launches: k_extend_primary + k_trace_dual + k_connect (SceneNearestHit / SceneAnyHit)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": traffic, "traffic_provenance": traffic_prov,
        "hbm": hbm, "l2": l2, "issue": issue,
        "algorithmic_bytes_per_launch": trav_b / max(trav_n, 1), "launch_ms": trav_ms / max(trav_n, 1), "launches": trav_n,
        "algorithmic_bytes_per_step": trav_b / steps, "traversal_ms_per_step": trav_ms / steps,
        "timed_pass": None if not serial_pass else {
            "ms_per_step": serial_pass[0] / max(serial_pass[1], 1), "steps": serial_pass[1],
            "what": "the kernel times of this object (and kernel_ms_per_step) come from a second timed pass over the same steps with the "
                    "library's per-family timers on; the timers make every wave run unsplit on one stream, because a kernel's duration is "
                    "only defined while kernels do not overlap.  `value` / `ms_per_step` of the line are the normal path, in which the two "
                    "halves of a wave share the GPU on two streams (a few per cent faster than this pass)"},
        "per_ray_nearest": {"n_inner": stats["n_inner"] / nr, "n_boxes": stats["n_boxes"] / nr, "n_leaf": stats["n_leaf"] / nr, "n_tri": stats["n_tri"] / nr, "n_switch": stats["n_switch"] / nr},
        "per_ray_any": {"n_inner": stats["n_inner_any"] / na, "n_leaf": stats["n_leaf_any"] / na, "n_tri": stats["n_tri_any"] / na, "n_switch": stats["n_switch_any"] / na},
        "note": "achieved / peak / frac: algorithmic bytes with the reference's record sizes (SURVEY 8(d): 64 B inner visit, 16 B leaf, 52 B "
                "triangle, 64 B switch) over the measured HBM copy peak; most of those bytes are served by L1 / L2, so this fraction may exceed 1",
    }


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
    port of the same algorithm is timed instead, on all host cores.  Exactly --steps timed steps after exactly
    --warmup warm-up steps, as in our arm; one step = the full 1080p frame at a bounded number of samples per
    pixel, sized from a one-sample calibration pass so that the whole run fits --ref-budget-s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from cadrays_b200.view import V3d_View
    from oracle.oracle_ffi import OracleScene
    desc = make_scene(args)
    host = V3d_View(host_only=True)
    desc.apply(host, with_target=False)
    blob = host.ExportBVH()
    host.Remove()
    cores = os.cpu_count() or 1
    w, h = args.width, args.height
    orc = OracleScene(blob)
    orc.configure(desc)
    acc = np.zeros((h, w, 4), dtype=np.float32)
    t0 = time.perf_counter()
    orc.render(w, h, 1, first_sample=1 << 20, accum=acc, nthreads=cores)       # calibration: one sample per pixel
    t_one = max(time.perf_counter() - t0, 1e-4)
    n_steps, n_warm = max(1, args.steps), max(0, args.warmup)
    spp = int(max(1, min(args.spp, args.ref_budget_s / (t_one * (n_steps + n_warm)))))
    for s_ in range(n_warm):
        acc[:] = 0
        orc.render(w, h, spp, first_sample=s_ * spp, accum=acc, nthreads=cores)
    t0 = time.perf_counter()
    for s_ in range(n_steps):
        acc[:] = 0
        orc.render(w, h, spp, first_sample=(n_warm + s_) * spp, accum=acc, nthreads=cores)
    dt = (time.perf_counter() - t0) / n_steps
    orc.close()
    v = (w * h * spp) / dt / 1e6
    sample = f"{w}x{h} x {spp} spp of the same view and scene per step (our arm: {args.spp} spp per step), {n_steps} steps after {n_warm} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps,
        "warmup": n_warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, desc),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle port (OpenMP over pixel rows); the OCCT/llvmpipe reference binary is not runnable here. The oracle "
                "walks the BVH the product's host builder exported (libcadrays_b200.so is mapped for that CPU-side build only)",
    }
    _OUT.write(json.dumps(line) + "\n"); _OUT.flush()


def _measure(view, desc, args, B, steps, warmup, rank, local, world, stream, accum, reduced, torch, dist, D, reduce=True):
    """Warm-up, the device-timed steps (CUDA events on the context's stream, max over ranks), then the same steps again
    with the instrumented kernels for the work counters.  With world > 1 the snapshot of every step's sums is all-reduced
    on a second stream while the next step traces."""
    comm = torch.cuda.Stream(device=torch.device("cuda", local)) if (world > 1 and reduce) else None
    reduce_done = torch.cuda.Event() if comm else None
    snapshot_ready = torch.cuda.Event() if comm else None

    def one_step(step_index):
        view.SetNextSample(D.step_sample_start(step_index, rank, world, B))
        view.RedrawAsync(B)
        if comm:
            stream.wait_event(reduce_done)           # no-op before the first record
            reduced.copy_(accum, non_blocking=True)
            snapshot_ready.record(stream)
            comm.wait_event(snapshot_ready)
            with torch.cuda.stream(comm):
                dist.all_reduce(reduced)
                reduce_done.record(comm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for s in range(warmup):
            one_step(s)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        view.ResetStats()
        ev0.record(stream)
        for s in range(steps):
            one_step(warmup + s)
        if comm:
            stream.wait_event(reduce_done)           # the timed region ends after the last all-reduce
        ev1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = ev0.elapsed_time(ev1)
        launches = view.LaunchCount()                # kernels of this library enqueued inside the timed region
        t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        # per-kernel accounting: the same steps once more with the library's per-family timers on.  A wave normally runs
        # as two halves on two streams (their kernels overlap, which is where 3 % of the throughput comes from), and a
        # kernel's duration is only defined while kernels do not overlap: with the timers on the library runs every wave
        # unsplit on one stream.  The roofline's kernel times and its own step time come from this pass.
        view.ResetStats()
        view.EnableTiming(True)
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record(stream)
        for s in range(steps):
            view.SetNextSample(D.step_sample_start(warmup + s, rank, world, B))
            view.RedrawAsync(B)
        ev3.record(stream)
        view.Sync()
        timing = view.Timing()
        timing["_serial_pass_ms"] = (ev2.elapsed_time(ev3), steps)
        timing["_launches"] = (0.0, launches)
        view.EnableTiming(False)
        # the same steps again with the instrumented kernels: work counters (untimed)
        view.EnableStats(True)
        view.ResetStats()
        for s in range(steps):
            view.SetNextSample(D.step_sample_start(warmup + s, rank, world, B))
            view.RedrawAsync(B)
        view.Sync()
        stats = view.Stats()
        view.EnableStats(False)
    return ms_total, timing, stats, clocks


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
    # host-side barrier (gloo): an NCCL barrier keeps a kernel spinning on the GPUs of the ranks that wait, and the
    # group e2e below has rank 0 render on every GPU while the other ranks wait
    host_group = dist.new_group(backend="gloo") if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def setup(desc, B):
        t0 = time.perf_counter()
        view = V3d_View(local)
        desc.apply(view)
        t_build = time.perf_counter() - t0
        p = desc.params
        p.SamplesPerBatch = B
        p.BvhWidth = args.bvh_width
        view.SetRenderingParams(p)
        view.Update()
        stream = torch.cuda.ExternalStream(view.Stream(), device=dev)
        accum = torch.zeros((desc.height, desc.width, 4), dtype=torch.float32, device=dev)
        reduced = torch.zeros_like(accum) if world > 1 else None
        torch.cuda.synchronize()
        view.BindAccum(accum.data_ptr(), accum.numel() * 4)
        return view, stream, accum, reduced, t_build

    desc = make_scene(args)
    W, H, B, depth = args.width, args.height, args.spp, args.depth
    view, stream, accum, reduced, t_build = setup(desc, B)
    ms_total, timing, stats, clocks = _measure(view, desc, args, B, args.steps, args.warmup, rank, local, world, stream, accum,
                                               reduced, torch, dist, D)
    serial_pass = timing.pop("_serial_pass_ms", None)
    launches_timed = timing.pop("_launches", (0.0, 0))[1]
    samples_per_step = W * H * B * world
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6
    rays = stats["rays_nearest"] + stats["rays_any"]
    mrays = rays * world / (ms_total * 1e-3) / 1e6

    # ---- strong scaling beside the weak line: the SAME total work per step (args.spp samples per pixel in all),
    # split over the ranks by sample index, summed with the same all-reduce
    strong = None
    if world > 1:
        Bs = max(1, B // world)
        p = desc.params
        p.SamplesPerBatch = Bs
        view.SetRenderingParams(p)
        view.BindAccum(accum.data_ptr(), accum.numel() * 4)
        ms_s, _tm, _, _ = _measure(view, desc, args, Bs, args.steps, args.warmup, rank, local, world, stream, accum, reduced, torch, dist, D)
        strong = {"scaling": "strong", "spp_per_step_total": Bs * world, "spp_per_step_per_gpu": Bs,
                  "value": W * H * Bs * world * args.steps / (ms_s * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_s / args.steps}
        p.SamplesPerBatch = B
        view.SetRenderingParams(p)
        view.BindAccum(accum.data_ptr(), accum.numel() * 4)

    # ---- e2e: the public call a user makes, host buffers, copies inside the timed region.  Every step restarts the
    # accumulation (SetCamera), renders B samples per pixel per rank, and ends with ONE combined tone-mapped RGB8 frame
    # in a pinned host buffer of rank 0; with N > 1 that includes the all-reduce of the N partial sums.
    ldr = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy()    # pinned host frame buffer
    cam = desc.camera
    if world == 1:
        view.BindAccum(None)

    def e2e_step(s):
        view.SetCamera(cam)                          # host -> device: camera / params (kernel arguments) + frame seeds
        view.SetNextSample(D.step_sample_start(s, rank, world, B))
        if world == 1:
            view.Redraw(B)                           # V3d_View::Redraw
            view.BufferDump(Graphic3d_BT_RGB, ldr)   # device -> host: tone-mapped RGB8 frame
        else:
            view.RedrawAsync(B)
            with torch.cuda.stream(stream):
                dist.all_reduce(accum)               # NCCL over NVLink, on the context's stream, in place (the step owns the buffer)
            if rank == 0:
                view.DumpFrom(accum.data_ptr(), ldr)                # display pass on the summed buffer + device -> host
            else:
                view.Sync()

    for s in range(2):
        e2e_step(s)
    e2e_steps = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_step(2 + s)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = W * H * B * world * e2e_steps / float(te.item()) / 1e6
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * B + 64, "d2h_bytes_per_step": W * H * 3,
           "steps": e2e_steps,
           "call": "SetCamera + Redraw(spp) + BufferDump(RGB8) per step, wall clock" if world == 1 else
                   "per step and rank: SetCamera + Redraw(spp) of the rank's sample block, NCCL all-reduce of the float4 sums, "
                   "then rank 0 runs the display pass on the combined buffer and copies ONE RGB8 frame to pinned host memory; wall clock, max over ranks"}

    line = None
    if rank == 0:
        trav_bytes, total_bytes = view.SceneBytes()
        extras = world == 1 and not args.no_extras
        roofline = build_roofline(args.workload, stats, timing, args.steps, trav_bytes, clocks, probe_dev=local if extras else None, serial_pass=serial_pass)
        kernel_ms = {k: v[0] / args.steps for k, v in timing.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, desc),
            "mrays_per_s": mrays, "rays_per_sample": rays / max(stats["samples"], 1),
            "roofline": roofline, "kernel_ms_per_step": kernel_ms,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches_timed),
            "scene_commit_s": t_build, "scene_bytes": {"traversal": trav_bytes, "total": total_bytes},
        }
        if strong:
            line["strong"] = strong
        if desc.env_source:
            line["config"]["environment"] = desc.env_source

    # ---- the reference's own cadence: one sample per pixel per Redraw(), then the frame is displayed
    # (AppViewer.cxx:1045-1047; the FPS the headless mode writes, main.cxx:218-227)
    if rank == 0 and world == 1 and not args.no_extras:
        p = desc.params
        p.SamplesPerBatch = 1
        view.SetRenderingParams(p)
        view.SetCamera(cam)
        for _ in range(10):
            view.Redraw(1); view.BufferDump(Graphic3d_BT_RGB, ldr)
        frames = 200
        t0 = time.perf_counter()
        for _ in range(frames):
            view.Redraw(1)                           # accumulation continues, as in the viewer
            view.BufferDump(Graphic3d_BT_RGB, ldr)
        dt1 = time.perf_counter() - t0
        line["e2e_1spp"] = {"frames_per_s": frames / dt1, "value": W * H * frames / dt1 / 1e6, "unit": UNIT, "ms_per_frame": dt1 / frames * 1e3,
                            "frames": frames, "h2d_bytes_per_step": 4, "d2h_bytes_per_step": W * H * 3,
                            "call": "Redraw(1) + BufferDump(RGB8) per frame, host buffer, wall clock -- the call pattern of AppViewer.cxx:1045-1047"}
    blob = view.ExportBVH() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    view.Remove()
    del accum, reduced
    torch.cuda.empty_cache()

    # ---- e2e through the call a single-process host makes (crt_group behind V3d_View: AppViewer.cxx:1047 Redraw,
    # :1259-1262 BufferDump): rank 0 alone drives all N GPUs -- scene built once on the host and uploaded to every
    # GPU, the step's samples dealt out over the GPUs, the partial sums exchanged over NVLink peer memory and
    # tone-mapped by the fused kernel, ONE combined RGB8 frame in a pinned host buffer.  The other ranks' GPUs are idle
    # (their contexts are gone) and their processes wait on a host-side barrier.
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                gv = V3d_View(devices=list(range(world)))
                tg0 = time.perf_counter()
                desc.apply(gv)
                t_group_commit = time.perf_counter() - tg0
                p = desc.params
                p.SamplesPerBatch = B
                gv.SetRenderingParams(p)
                gv.Update()

                def group_run(spp_total, n_steps):
                    for _ in range(2):
                        gv.SetCamera(cam); gv.Redraw(spp_total); gv.BufferDump(Graphic3d_BT_RGB, ldr)
                    t0 = time.perf_counter()
                    for _ in range(n_steps):
                        gv.SetCamera(cam)                         # restarts the accumulation on every GPU
                        gv.Redraw(spp_total)                      # one call, N GPUs
                        gv.BufferDump(Graphic3d_BT_RGB, ldr)      # exchange + Display + device -> host
                    return time.perf_counter() - t0
                dtw = group_run(B * world, e2e_steps)
                dts = group_run(B, e2e_steps)
                info = gv.GroupInfo()
                line["e2e_ranks"] = line["e2e"]
                line["e2e"] = {
                    "value": W * H * B * world * e2e_steps / dtw / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * B * world + 64 * world,
                    "d2h_bytes_per_step": W * H * 3, "steps": e2e_steps, "scaling": "weak",
                    "call": "ONE host process, crt_group over all N GPUs: SetCamera + Redraw(spp x N) + BufferDump(RGB8) per step "
                            "(samples dealt out over the GPUs, partial sums read over NVLink peer memory by the fused reduce + Display "
                            "kernel, one combined frame in a pinned host buffer); wall clock",
                    "exchange": "nccl" if info["nccl"] else "fused peer-memory kernel", "exchange_display_ms": info["reduce_ms"],
                    "group_commit_s": t_group_commit,
                    "strong": {"scaling": "strong", "spp_per_step_total": B, "value": W * H * B * e2e_steps / dts / 1e6, "unit": UNIT,
                               "ms_per_step": dts / e2e_steps * 1e3},
                }
                gv.Remove()
            except Exception as ex:                      # the per-rank e2e above stays the reported number
                line["e2e_group_error"] = repr(ex)
        dist.barrier(group=host_group)

    # ---- second block: the HBM-resident case, config C5 flattened (10.5 M unique triangles, 1.3 GB of scene data)
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "assembly":
        import copy
        a5 = copy.copy(args)
        a5.workload, a5.depth = "instanced_flat", 8
        d5 = make_scene(a5)
        v5, st5, acc5, red5, tb5 = setup(d5, B)
        n5 = max(3, args.steps // 2)
        ms5, tm5, s5, ck5 = _measure(v5, d5, a5, B, n5, min(args.warmup, 3), 0, local, 1, st5, acc5, red5, torch, dist, D)
        sp5 = tm5.pop("_serial_pass_ms", None)
        tm5.pop("_launches", None)
        tb, tt = v5.SceneBytes()
        r5 = s5["rays_nearest"] + s5["rays_any"]
        line["c5_flattened"] = {
            "config": {"workload": f"C5 flattened: {len(d5.instances)} objects, {d5.n_triangles()} unique triangles, 1 directional light, "
                                   f"{W}x{H}, depth 8", "spp_per_step_per_gpu": B},
            "value": W * H * B * n5 / (ms5 * 1e-3) / 1e6, "unit": UNIT, "steps": n5, "ms_per_step": ms5 / n5,
            "mrays_per_s": r5 / (ms5 * 1e-3) / 1e6, "scene_bytes": {"traversal": tb, "total": tt}, "scene_commit_s": tb5,
            "roofline": build_roofline("instanced_flat", s5, tm5, n5, tb, ck5, probe_dev=local, serial_pass=sp5),
            "kernel_ms_per_step": {k: v[0] / n5 for k, v in tm5.items()}, "clocks": ck5,
        }
        v5.Remove()
        del acc5
        torch.cuda.empty_cache()

    if rank == 0:
        if blob is not None:
            v, dtc, sample = cpu_leg(args, desc, blob, os.cpu_count() or 1, 3)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
        _OUT.write(json.dumps(line) + "\n"); _OUT.flush()
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
