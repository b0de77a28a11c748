"""Regression harness with the contract of the reference's testing/CADRays_Testing.py (the only test driver it has):

  python -m cadrays_b200.regress -i <folder with .tcl scripts> -f <frames> -m <template folder> [-o <output folder>] [-d <percent>]
  python -m cadrays_b200.regress -o <output folder> -m <template folder> -u

Run mode renders every script of the folder for N frames with the headless runner (`cadrays_b200.run`, the counterpart
of `CADRays.exe script.tcl N`), collects Output_<script>_<N>.png / .txt into <output>/<dd_mm_YYYY HH_MM_SS>/ and
writes Result.html there: frame rate of every script next to the template's (marked when it moved by more than -d
percent, default 2), and output image / template image / difference mask (white where any channel differs --
rendering is deterministic here, so the mask is empty unless something changed).  Update mode (-u) makes the newest
run the template: copies its images as <script>.png and its frame rates as Result.html into the template folder.

The reference spawns `CADRays.exe`; here the renderer is called through `runner(script_path, frames, out_dir)`,
`cadrays_b200.run.main` by default, so the bookkeeping can be exercised without a GPU.
"""
from __future__ import annotations

import argparse
import html
import os
import re
import shutil
import sys
from datetime import datetime
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import imageio

STAMP = "%d_%m_%Y %H_%M_%S"
_STAMP_RE = re.compile(r"^\d{1,2}_\d{1,2}_\d{4} \d{1,2}_\d{1,2}_\d{1,2}$")
_RATE_RE = re.compile(r"<!-- rate (?P<name>.*?) = (?P<fps>[-+0-9.eE]+) -->")


def default_runner(script: str, frames: int, out_dir: str) -> None:
    from . import run
    rc = run.main([script, str(frames), "--out", out_dir])
    if rc != 0:
        raise RuntimeError(f"{script}: runner returned {rc}")


def read_rates(result_html: str) -> Dict[str, float]:
    """Frame rates recorded in a Result.html written by write_report (machine-readable comments)."""
    if not os.path.isfile(result_html):
        return {}
    with open(result_html, encoding="utf-8") as f:
        return {m.group("name"): float(m.group("fps")) for m in _RATE_RE.finditer(f.read())}


def difference_mask(a: np.ndarray, b: np.ndarray) -> Optional[np.ndarray]:
    """White where the two RGB images differ in any channel (None when the sizes differ)."""
    if a.shape != b.shape:
        return None
    m = np.any(a != b, axis=2)
    return np.repeat((m * 255).astype(np.uint8)[..., None], 3, axis=2)


def write_report(path: str, when: datetime, rates: List[Tuple[str, float]], template: Dict[str, float], max_diff: float,
                 images: List[Tuple[str, str, str, str, Optional[int]]]) -> None:
    out = ["<html><head><meta charset='utf-8'><title>Result</title></head><body>", f"<h1>{when.strftime('%d/%m/%Y %H:%M:%S')}</h1>", "<ol>"]
    for name, fps in rates:
        line = f"Framerate = {fps:.3f} fps"
        style = ""
        if name in template and template[name] > 0:
            change = (fps / template[name] - 1.0) * 100.0
            line += f" (prev = {template[name]:.3f}) [{change:+.4f}%]"
            if abs(change) > max_diff:
                style = " style='background-color:%s'" % ("green" if change > 0 else "red")
        out.append(f"<li><p><strong>File {html.escape(name)}</strong></p><ul><li><p><strong><span{style}>{line}</span></strong></p>"
                   f"<!-- rate {html.escape(name)} = {fps!r} --></li></ul></li>")
    for name, result_png, model_png, diff_png, n_diff in images:
        out.append(f"<li><p><strong>File {html.escape(name)}</strong></p>")
        if result_png and model_png:
            verdict = "sizes differ" if n_diff is None else ("identical" if n_diff == 0 else f"{n_diff} pixels differ")
            out.append(f"<p>{verdict}</p><table><tr><th>Output result</th><th>Model result</th><th>Difference</th></tr><tr>"
                       + "".join(f"<td><img src='file://{html.escape(p)}' width='100%'></td>" for p in (result_png, model_png, diff_png) if p)
                       + "</tr></table>")
        elif result_png:
            out.append(f"<table><tr><th>Output result</th></tr><tr><td><img src='file://{html.escape(result_png)}' width='100%'></td></tr></table>")
        out.append("</li>")
    out += ["</ol>", "</body></html>"]
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(out) + "\n")


def run_folder(scripts_dir: str, frames: int, out_root: str, model_dir: str, max_diff: float = 2.0,
               runner: Callable[[str, int, str], None] = default_runner, now: Optional[datetime] = None) -> str:
    """Renders every .tcl of scripts_dir, compares with the template folder, returns the run folder."""
    when = now or datetime.now()
    run_dir = os.path.join(out_root, when.strftime(STAMP))
    os.makedirs(run_dir, exist_ok=True)
    scripts = sorted(f for f in os.listdir(scripts_dir) if f.lower().endswith(".tcl") and os.path.isfile(os.path.join(scripts_dir, f)))
    template = read_rates(os.path.join(model_dir, "Result.html"))
    rates: List[Tuple[str, float]] = []
    images: List[Tuple[str, str, str, str, Optional[int]]] = []
    for script in scripts:
        stem = os.path.splitext(script)[0]
        runner(os.path.join(scripts_dir, script), frames, run_dir)
        txt = os.path.join(run_dir, f"Output_{stem}_{frames}.txt")
        png = os.path.join(run_dir, f"Output_{stem}_{frames}.png")
        if os.path.isfile(txt):
            with open(txt) as f:
                rates.append((script, float(f.readline().strip() or 0.0)))
            os.remove(txt)               # the reference removes the .txt files once their number is in the report
        result_png = png if os.path.isfile(png) else ""
        model_png = os.path.join(model_dir, stem + ".png")
        model_png = model_png if os.path.isfile(model_png) else ""
        diff_png, n_diff = "", None
        if result_png and model_png:
            mask = difference_mask(imageio.read_png_rgb8(result_png), imageio.read_png_rgb8(model_png))
            if mask is not None:
                diff_png = os.path.join(run_dir, f"Diff_{stem}.png")
                imageio.write_png(diff_png, mask[::-1])
                n_diff = int(np.count_nonzero(mask[..., 0]))
        images.append((script, result_png, model_png, diff_png, n_diff))
    write_report(os.path.join(run_dir, "Result.html"), when, rates, template, max_diff, images)
    return run_dir


def newest_run(out_root: str) -> Optional[str]:
    runs = [d for d in os.listdir(out_root) if _STAMP_RE.match(d) and os.path.isdir(os.path.join(out_root, d))]
    if not runs:
        return None
    return os.path.join(out_root, max(runs, key=lambda d: datetime.strptime(d, STAMP)))


def update_template(out_root: str, model_dir: str) -> str:
    """-u: the newest run under out_root becomes the template."""
    run_dir = newest_run(out_root)
    if run_dir is None:
        raise FileNotFoundError("no results found")
    rates = read_rates(os.path.join(run_dir, "Result.html"))
    os.makedirs(model_dir, exist_ok=True)
    write_report(os.path.join(model_dir, "Result.html"), datetime.strptime(os.path.basename(run_dir), STAMP),
                 sorted(rates.items()), {}, 0.0, [])
    for f in os.listdir(run_dir):
        m = re.match(r"^Output_(.*)_(\d+)\.png$", f)
        if m:
            shutil.copyfile(os.path.join(run_dir, f), os.path.join(model_dir, m.group(1) + ".png"))
    return run_dir


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("-i", dest="scripts", default="", help="folder with the .tcl scripts")
    ap.add_argument("-f", dest="frames", type=int, default=100, help="frames per script")
    ap.add_argument("-d", dest="max_diff", type=float, default=2.0, help="frame-rate change (percent) that gets highlighted")
    ap.add_argument("-o", dest="out", default="", help="output folder (default: the scripts folder)")
    ap.add_argument("-m", dest="model", default="", help="template folder")
    ap.add_argument("-u", dest="update", action="store_true", help="make the newest run the template")
    args = ap.parse_args(argv)
    if not args.model or (not args.update and not os.path.isdir(args.model)):
        print("Path to the folder with results for comparing is incorrect", file=sys.stderr)
        return 2
    if args.update:
        if not args.out or not os.path.isdir(args.out):
            print("Path to output folder is incorrect", file=sys.stderr)
            return 2
        try:
            print("template updated from", update_template(args.out, args.model))
        except FileNotFoundError as e:
            print(str(e), file=sys.stderr)
            return 2
        return 0
    if not args.scripts or not os.path.isdir(args.scripts):
        print("Path to scripts folder is incorrect", file=sys.stderr)
        return 2
    out = args.out or args.scripts
    if not os.path.isdir(out):
        print("Path to output folder is incorrect", file=sys.stderr)
        return 2
    print("report:", os.path.join(run_folder(args.scripts, args.frames, out, args.model, args.max_diff), "Result.html"))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
