"""Headless run, the counterpart of `CADRays.exe script.tcl N` (src/Launcher/main.cxx:164-229,
src/Launcher/AppViewer.cxx:1059-1071,1255-1264): evaluates the script, renders N frames (one sample
per pixel per Redraw(), like the reference's GUI loop), writes Output_<script>_<N>.png (BufferDump RGB)
and Output_<script>_<N>.txt (average frames per second) -- the two files testing/CADRays_Testing.py reads.

  python -m cadrays_b200.run script.tcl N [--size WxH] [--out DIR] [--hdr] [--device 0] [--spp-per-redraw K]

A script that renders by itself (`vfps N` followed by `vdump file`, as data/other/preview.tcl does for every named
material) gets each dump rendered when the command is reached: N frames (capped by --max-dump-frames), written as
<out>/<basename of file>.
"""
from __future__ import annotations

import argparse
import os
import time

from . import imageio, tcl
from .view import Graphic3d_BT_RGB, Graphic3d_BT_RGB_RayTraceHdrLeft, V3d_View


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("frames", type=int)
    ap.add_argument("--size", default="1900x1000")       # main.cxx:103 creates the viewer at 1900x1000
    ap.add_argument("--out", default=".")
    ap.add_argument("--hdr", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--spp-per-redraw", type=int, default=1)
    ap.add_argument("--max-dump-frames", type=int, default=8000)
    args = ap.parse_args(argv)
    w, h = (int(v) for v in args.size.lower().split("x"))
    view = V3d_View(args.device)
    sess = tcl.DrawSession(w, h, root=os.path.dirname(os.path.abspath(args.script)))
    sess.size_fixed = True

    def on_dump(session, path, frames):
        d = session.scene()
        d.apply(view)
        view.Redraw(max(1, min(frames or 1, args.max_dump_frames)))
        target = os.path.join(args.out, os.path.basename(path.replace("\\", "/")))
        imageio.write_png(os.path.splitext(target)[0] + ".png", view.BufferDump(Graphic3d_BT_RGB))

    sess.on_dump = on_dump
    with open(args.script, "r", encoding="utf-8", errors="replace") as f:
        sess.eval(f.read())
    desc = sess.scene()
    desc.apply(view)
    t0 = time.perf_counter()
    frames = 0
    while frames < args.frames:
        k = min(args.spp_per_redraw, args.frames - frames)
        view.Redraw(k)
        frames += k
    dt = time.perf_counter() - t0
    name = os.path.splitext(os.path.basename(args.script))[0]
    stem = os.path.join(args.out, f"Output_{name}_{args.frames}")
    imageio.write_png(stem + ".png", view.BufferDump(Graphic3d_BT_RGB))
    if args.hdr:
        imageio.write_hdr(stem + ".hdr", view.BufferDump(Graphic3d_BT_RGB_RayTraceHdrLeft))
    with open(stem + ".txt", "w") as f:
        f.write(f"{frames / dt:.3f}\n")
    if sess.unknown:
        print("ignored commands:", sorted(set(sess.unknown)))
    view.Remove()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
