"""Offline SIMT cost model of traversal control-flow strategies (analysis tooling, CPU only).

Lives under tests/ because it drives the CPU oracle (`orc_trace_events`), and the oracle is test infrastructure: nothing
outside tests/, smoke() and the CPU legs of bench.py may touch it.

Takes real per-ray traversal event strings from the oracle (1 = inner node, 2 = instance switch,
3+k = leaf with k triangles) for a scene, groups rays into 32-lane warps with per-lane refill, and
counts warp-level instruction issue for several loop structures.  Instruction costs per step come
from the SASS of the current kernels.  Used to decide what is worth building before spending GPU time.

  python tests/analysis/simt_model.py [--rays 65536] [--scene assembly]
"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))

C_N, C_S, C_L0, C_T, C_R = 50, 100, 15, 45, 40


def get_events(desc, org, d, tmax=None, any_hit=False, cap=512):
    from cadrays_b200.view import V3d_View
    from oracle import oracle_ffi
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    o = oracle_ffi.OracleScene(v.ExportBVH())
    L = oracle_ffi.lib()
    L.orc_trace_events.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32,
                                   C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int32)]
    n = org.shape[0]
    ev = np.zeros((n, cap), np.uint8)
    lens = np.zeros(n, np.int32)
    fp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
    L.orc_trace_events(o._h, fp(org), fp(d), fp(tmax), n, int(any_hit), ev.ctypes.data_as(C.POINTER(C.c_uint8)), cap,
                       lens.ctypes.data_as(C.POINTER(C.c_int32)))
    prim, inst, t, u, vv = o.trace(org, d, tmax, any_hit=any_hit)
    return [ev[i, :lens[i]].tolist() for i in range(n)], prim, t


def secondary_rays(desc, n, seed=1):
    """Camera rays -> hit points -> random directions in the hemisphere facing the camera (diffuse-bounce proxy)."""
    from cadrays_b200.view import V3d_View
    from oracle import oracle_ffi
    g = np.random.default_rng(seed)
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    o = oracle_ffi.OracleScene(v.ExportBVH())
    o.configure(desc)
    L = oracle_ffi.lib()
    w, h = desc.width, desc.height
    side = int(np.sqrt(n * 2.5))
    xs = (np.arange(side) + 0.5) / side
    org = np.zeros((side * side, 3), np.float32); d = np.zeros_like(org)
    oo = (C.c_float * 3)(); dd = (C.c_float * 3)()
    k = 0
    # tile order 8x4 like the kernel: iterate tiles then lanes
    for ty in range(0, side, 4):
        for tx in range(0, side, 8):
            for ly in range(4):
                for lx in range(8):
                    x, y = tx + lx, ty + ly
                    if x >= side or y >= side:
                        continue
                    L.orc_camera_ray(o._h, float(xs[x]), float(xs[y]), 0.0, 0.0, oo, dd)
                    org[k] = oo[:]; d[k] = dd[:]; k += 1
    org, d = org[:k], d[:k]
    prim, inst, t, u, vv = o.trace(org, d)
    hit = prim >= 0
    p = org[hit] + d[hit] * t[hit, None]
    r = g.normal(size=p.shape); r /= np.linalg.norm(r, axis=1, keepdims=True)
    flip = np.sum(r * (-d[hit]), axis=1) < 0
    r[flip] = -r[flip]
    eps = 1e-3
    sec_o = (p + r * eps - d[hit] * eps).astype(np.float32)
    return (org, d), (sec_o[:n], r[:n].astype(np.float32))


def simulate(events, strategy, threshold=1):
    """Returns (warp instruction issues, useful lane instructions)."""
    n = len(events)
    pos_ray = 0
    total_issue = 0
    useful = 0
    for e in events:
        for c in e:
            useful += C_N if c == 1 else (C_S if c == 2 else C_L0 + C_T * (c - 3))
    # one persistent warp per 32 * 64 rays is enough to see the steady state; process warps sequentially
    chunk = 32 * 64
    for base in range(0, n, chunk):
        pool = list(range(base, min(base + chunk, n)))
        pool.reverse()
        lanes = [None] * 32          # (event list, index)
        while True:
            # refill
            need = [i for i in range(32) if lanes[i] is None]
            if need and pool:
                total_issue += C_R
                for i in need:
                    if pool:
                        lanes[i] = [events[pool.pop()], 0]
            for i in range(32):
                if lanes[i] is not None and lanes[i][1] >= len(lanes[i][0]):
                    lanes[i] = None
            if all(l is None for l in lanes):
                if not pool:
                    break
                continue
            nxt = lambda l: l[0][l[1]] if l is not None and l[1] < len(l[0]) else 0
            if strategy in ("while-while", "inloop-switch"):
                # inner loop
                while True:
                    kinds = [nxt(l) for l in lanes]
                    n_inner = sum(1 for k in kinds if k == 1)
                    n_sw = sum(1 for k in kinds if k == 2) if strategy == "inloop-switch" else 0
                    if n_inner + n_sw == 0:
                        break
                    if n_inner < threshold and any(k > (2 if strategy == "inloop-switch" else 1) for k in kinds):
                        break
                    total_issue += (C_N if n_inner else 0) + (C_S if n_sw else 0) + (4 if threshold > 1 else 0)
                    for l, k in zip(lanes, kinds):
                        if k == 1 or (k == 2 and strategy == "inloop-switch"):
                            l[1] += 1
                kinds = [nxt(l) for l in lanes]
                if any(k == 2 for k in kinds):
                    total_issue += C_S
                ks = [k - 3 for k in kinds if k >= 3]
                if ks:
                    total_issue += C_L0 + C_T * max(ks)
                for l, k in zip(lanes, kinds):
                    if k >= 2:
                        l[1] += 1
            elif strategy == "if-if":
                kinds = [nxt(l) for l in lanes]
                if any(k == 1 for k in kinds):
                    total_issue += C_N
                if any(k == 2 for k in kinds):
                    total_issue += C_S
                ks = [k - 3 for k in kinds if k >= 3]
                if ks:
                    total_issue += C_L0 + C_T * max(ks)
                for l, k in zip(lanes, kinds):
                    if k:
                        l[1] += 1
            elif strategy == "tri-step":
                # leaves are processed one triangle per iteration, inside the same loop as inner nodes
                kinds = [nxt(l) for l in lanes]
                if any(k == 1 for k in kinds):
                    total_issue += C_N
                if any(k == 2 for k in kinds):
                    total_issue += C_S
                if any(k >= 3 for k in kinds):
                    total_issue += C_T + 5
                for l, k in zip(lanes, kinds):
                    if k in (1, 2) or k == 3 or k == 4:
                        l[1] += 1
                    elif k > 4:
                        l[0] = list(l[0]); l[0][l[1]] = k - 1
            for i in range(32):
                if lanes[i] is not None and lanes[i][1] >= len(lanes[i][0]):
                    lanes[i] = None
    return total_issue, useful


def simulate_phase_pool(events, pool_size=64, gather=40, exit_lanes=12, leaf_pairs=True):
    """A warp owns `pool_size` rays whose state lives in shared memory; it repeatedly picks up to 32 rays that are in the
    same phase (node walk / leaf / instance switch), gathers them into registers (`gather` instructions each way), and
    runs that phase with (nearly) full lanes; the node phase is left once fewer than `exit_lanes` lanes still walk.
    Leaf phase: one (ray, triangle) pair per lane and round when leaf_pairs, else one leaf per lane.  Returns
    (warp instruction issues, useful lane instructions) like simulate()."""
    n = len(events)
    useful = 0
    for e in events:
        for c in e:
            useful += C_N if c == 1 else (C_S if c == 2 else C_L0 + C_T * (c - 3))
    total = 0
    chunk = pool_size * 48
    for base in range(0, n, chunk):
        queue = list(range(base, min(base + chunk, n)))
        queue.reverse()
        pool = []                    # [events, index]
        while True:
            pool = [r for r in pool if r[1] < len(r[0])]
            if len(pool) < pool_size and queue:
                total += C_R
                while len(pool) < pool_size and queue:
                    pool.append([events[queue.pop()], 0])
                pool = [r for r in pool if r[1] < len(r[0])]
            if not pool:
                if not queue:
                    break
                continue
            walk = [r for r in pool if r[0][r[1]] == 1]
            sw = [r for r in pool if r[0][r[1]] == 2]
            leaf = [r for r in pool if r[0][r[1]] >= 3]
            # pick the phase with the most waiting rays (ties: walk)
            if len(walk) >= max(len(sw), len(leaf)) and walk:
                sel = walk[:32]
                total += 2 * gather
                first = True
                while True:
                    act = [r for r in sel if r[1] < len(r[0]) and r[0][r[1]] == 1]
                    if not act or (not first and len(act) < exit_lanes and len(pool) > len(sel)):
                        break
                    first = False
                    total += C_N
                    for r in act:
                        r[1] += 1
            elif len(leaf) >= len(sw) and leaf:
                sel = leaf[:32]
                total += 2 * gather
                if leaf_pairs:
                    pairs = sum(r[0][r[1]] - 3 for r in sel)
                    total += C_L0 + (C_T + 10) * max(1, (pairs + 31) // 32)
                else:
                    total += C_L0 + C_T * max(r[0][r[1]] - 3 for r in sel)
                for r in sel:
                    r[1] += 1
            else:
                sel = sw[:32]
                total += 2 * gather + C_S
                for r in sel:
                    r[1] += 1
    return total, useful


def simulate_lane_smt(events, k=2, gather=40, exit_lanes=12):
    """Every lane owns k rays (state outside the registers, stacks stay thread-local); the warp votes for a phase and each
    lane works on one of ITS rays that is in that phase, if it has one.  No ray ever changes lanes."""
    n = len(events)
    useful = 0
    for e in events:
        for c in e:
            useful += C_N if c == 1 else (C_S if c == 2 else C_L0 + C_T * (c - 3))
    total = 0
    chunk = 32 * k * 24
    for base in range(0, n, chunk):
        queue = list(range(base, min(base + chunk, n)))
        queue.reverse()
        lanes = [[None] * k for _ in range(32)]
        def phase(r):
            if r is None or r[1] >= len(r[0]): return 0
            c = r[0][r[1]]
            return 1 if c == 1 else (2 if c == 2 else 3)
        while True:
            refilled = False
            for l in range(32):
                for j in range(k):
                    if phase(lanes[l][j]) == 0:
                        lanes[l][j] = None
                        if queue:
                            lanes[l][j] = [events[queue.pop()], 0]; refilled = True
            if refilled:
                total += C_R
            counts = [0, 0, 0, 0]
            for l in range(32):
                ps = {phase(r) for r in lanes[l]}
                for p in (1, 2, 3):
                    if p in ps: counts[p] += 1
            if sum(counts) == 0:
                if not queue: break
                continue
            p = max((1, 3, 2), key=lambda x: counts[x])
            sel = []
            for l in range(32):
                for r in lanes[l]:
                    if phase(r) == p:
                        sel.append(r); break
            total += 2 * gather
            if p == 1:
                first = True
                while True:
                    act = [r for r in sel if phase(r) == 1]
                    if not act or (not first and len(act) < exit_lanes):
                        break
                    first = False
                    total += C_N
                    for r in act: r[1] += 1
            elif p == 3:
                total += C_L0 + C_T * max(r[0][r[1]] - 3 for r in sel)
                for r in sel: r[1] += 1
            else:
                total += C_S
                for r in sel: r[1] += 1
    return total, useful


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=32768)
    ap.add_argument("--scene", default="assembly")
    args = ap.parse_args()
    from cadrays_b200 import scenes
    desc = scenes.assembly(width=512, height=512) if args.scene == "assembly" else scenes.cornell_box(512, 512)
    (po, pd), (so, sd) = secondary_rays(desc, args.rays)
    for name, (o, d) in (("primary (coherent)", (po[:args.rays], pd[:args.rays])), ("secondary (incoherent)", (so, sd))):
        ev, prim, t = get_events(desc, o, d)
        steps = np.array([len(e) for e in ev])
        print(f"== {name}: {len(ev)} rays, mean steps {steps.mean():.1f}, max {steps.max()}, hit {np.mean(prim >= 0):.2f}")
        for strat, thr in (("while-while", 1), ("while-while", 12), ("while-while", 20), ("inloop-switch", 1), ("if-if", 1), ("tri-step", 1)):
            issue, useful = simulate(ev, strat, thr)
            print(f"   {strat:14s} thr={thr:2d}: warp issues {issue/len(ev):8.1f} per ray, lane efficiency {useful / (32 * issue):.3f}")
        for ps, g, ex in ((64, 40, 12), (64, 40, 20), (96, 40, 16), (128, 40, 16), (64, 70, 16), (64, 20, 16)):
            issue, useful = simulate_phase_pool(ev, ps, g, ex)
            print(f"   phase-pool {ps:3d} rays, gather {g}, exit<{ex}: warp issues {issue/len(ev):8.1f} per ray, lane efficiency {useful / (32 * issue):.3f}")
        for k, g, ex in ((2, 40, 12), (2, 40, 20), (3, 40, 16), (4, 40, 16), (2, 25, 16), (2, 60, 16)):
            issue, useful = simulate_lane_smt(ev, k, g, ex)
            print(f"   lane-smt {k} rays per lane, gather {g}, exit<{ex}: warp issues {issue/len(ev):8.1f} per ray, lane efficiency {useful / (32 * issue):.3f}")
        # sorted variant: group rays by origin cell + direction octant before forming warps
        lo, hi = o.min(0), o.max(0)
        cell = np.clip(((o - lo) / np.maximum(hi - lo, 1e-9) * 16).astype(int), 0, 15)
        key = (((d[:, 0] < 0) * 4 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0)) << 12) | (cell[:, 0] << 8) | (cell[:, 1] << 4) | cell[:, 2]
        order = np.argsort(key, kind="stable")
        ev_sorted = [ev[i] for i in order]
        issue, useful = simulate(ev_sorted, "while-while", 1)
        print(f"   {'sorted + w-w':14s}        : warp issues {issue/len(ev):8.1f} per ray, lane efficiency {useful / (32 * issue):.3f}")


if __name__ == "__main__":
    main()
