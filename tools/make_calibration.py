"""Turns the ncu launch lists of tools/profile_round2.sh into the two small files bench.py reads for its roofline:

  profiles/extend_traffic.json      dram bytes per traversal launch (roofline.traffic), per workload
  profiles/issue_calibration.json   issue-slot utilisation, lanes per instruction, L1 data-pipe utilisation and the
                                    measured thread instructions per unit of the SURVEY 8(d) instruction formula

usage: python tools/make_calibration.py <tag> <commit> [--stats bench.json]
  reads gpurun_out/<tag>_launches.csv (C2) and gpurun_out/<tag>_launches_c5flat.csv (C5 flattened): the launches of
  `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras [--workload instanced_flat]` under
  `ncu --clock-control none` with CRT_PIPELINE=0.  The last complete step in the list is used.  Time-weighted
  averages over the traversal launches of that step."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from launch_table import load

REPO = Path(__file__).resolve().parent.parent
TRAV = ("k_extend", "k_trace_dual", "k_connect")


def f(v, n):
    try:
        return float(v.get(n, "").replace(",", ""))
    except ValueError:
        return float("nan")


def one(path, formula_units_per_step=None):
    data = list(load(path).items())
    # the launch lists are taken with CRT_PIPELINE=0 (a wave unsplit on one stream; ncu serialises kernels anyway), so a
    # step is the run of launches from a k_extend_primary to the next k_resolve: the last complete one is used
    starts = [i for i, (k, _) in enumerate(data) if "k_extend_primary" in k[1]]
    ends = [i for i, (k, _) in enumerate(data) if "k_resolve" in k[1]]
    if not starts or not ends or ends[-1] < starts[0]:
        raise SystemExit(f"{path}: no complete step")
    b = ends[-1] + 1
    a = max(i for i in starts if i < ends[-1])
    parts = 1
    step = data[a:b]
    trav = [(k, v) for k, v in step if any(t in k[1] for t in TRAV)]
    t_all = sum(f(v, "gpu__time_duration.sum") for _, v in step)
    t_tr = sum(f(v, "gpu__time_duration.sum") for _, v in trav)
    w = lambda name: sum(f(v, name) * f(v, "gpu__time_duration.sum") for _, v in trav) / t_tr
    dram = sum(f(v, "dram__bytes_read.sum") + f(v, "dram__bytes_write.sum") for _, v in trav)
    winst = sum(f(v, "smsp__inst_executed.sum") for _, v in trav)
    tinst = sum(f(v, "smsp__thread_inst_executed.sum") for _, v in trav)
    return {
        "wave_parts": parts, "launches_in_step": len(step), "traversal_launches": len(trav),
        "traversal_ms_ncu": t_tr / 1e6, "step_ms_ncu": t_all / 1e6, "traversal_share_ncu": t_tr / t_all,
        "dram_bytes_per_step": dram, "dram_bytes_per_launch": dram / len(trav),
        "issue_active_pct": w("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": tinst / winst,
        "l1_data_pipe_pct": w("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "l2_pct": w("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": w("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": w("lts__t_sector_hit_rate.pct"),
        "warp_inst": winst, "thread_inst": tinst,
    }


def main():
    tag, commit = sys.argv[1], sys.argv[2]
    stats = {}
    if "--stats" in sys.argv:
        # formula units (45 n_inner + 55 n_tri + 60 n_switch + 30 per ray) of one step, from a bench line of the same build
        line = json.loads([l for l in open(sys.argv[sys.argv.index("--stats") + 1]) if l.startswith("{")][-1])
        for wl, blk in (("assembly", line), ("instanced_flat", line.get("c5_flattened") or {})):
            r = (blk.get("roofline") or {})
            if r.get("issue"):
                rays = blk["mrays_per_s"] * 1e6 * blk["ms_per_step"] * 1e-3
                k = r["issue"].get("calibration") or {}
                stats[wl] = r["issue"]["thread_inst_per_ray"] / float(k.get("thread_inst_per_formula_unit", 1.0)) * rays
    cmd = "ncu --metrics <tools/profile_round2.sh list> --clock-control none -c 80 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
    traffic, cal = {}, {}
    for wl, name in (("assembly", f"{tag}_launches.csv"), ("instanced_flat", f"{tag}_launches_c5flat.csv")):
        p = REPO / "gpurun_out" / name
        if not p.exists():
            continue
        o = one(p)
        src = f"profiles/{name.replace(tag, 'r02')}"
        traffic[wl] = {"dram_bytes_per_launch": o["dram_bytes_per_launch"], "dram_bytes_per_step": o["dram_bytes_per_step"],
                       "launches": o["traversal_launches"], "round": 2, "commit": commit,
                       "command": cmd + ("" if wl == "assembly" else " --workload instanced_flat"),
                       "source": f"dram__bytes_read.sum + dram__bytes_write.sum over the traversal launches of the last step in {src}"}
        e = {k: o[k] for k in ("issue_active_pct", "lanes_per_inst", "l1_data_pipe_pct", "l2_pct", "l1_hit_pct", "l2_hit_pct", "wave_parts",
                                "traversal_share_ncu", "traversal_ms_ncu", "step_ms_ncu")}
        e["thread_inst_per_step_ncu"] = o["thread_inst"]
        if wl in stats and stats[wl] > 0:
            e["thread_inst_per_formula_unit"] = o["thread_inst"] / stats[wl]
        e.update({"round": 2, "commit": commit, "source": src, "averaging": "time-weighted over the traversal launches of one step"})
        cal[wl] = e
    json.dump(traffic, open(REPO / "profiles" / "extend_traffic.json", "w"), indent=1)
    json.dump(cal, open(REPO / "profiles" / "issue_calibration.json", "w"), indent=1)
    print(json.dumps(cal, indent=1))


if __name__ == "__main__":
    main()
