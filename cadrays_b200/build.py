"""In-tree build of libcadrays_b200.so (nvcc, sm_100a) -- no JIT cache, the .so travels with the tree.

Usage: python -m cadrays_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libcadrays_b200.so"
PROBE_PATH = PKG_DIR / "libcrt_probe.so"      # memory-hierarchy probe for the bench's roofline denominators

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # arithmetic contract: no implicit contraction, IEEE div/sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-fopenmp,-Wall",
    "-shared",
]


def _sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [d for d in CSRC.glob("*") if d.is_file()] + [PKG_DIR.parent / "include" / "cadrays_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_probe(force: bool = False) -> Path:
    """libcrt_probe.so: L2 / HBM bandwidth probe (csrc/probe/mem_probe.cu), sm_100a, in-tree."""
    src = CSRC / "probe" / "mem_probe.cu"
    if not force and PROBE_PATH.exists() and PROBE_PATH.stat().st_mtime >= src.stat().st_mtime:
        return PROBE_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-o", str(PROBE_PATH), str(src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (probe):\n" + res.stdout + res.stderr)
    return PROBE_PATH


def build_library(force: bool = False, verbose: bool = False, defines=(), out: Path | None = None) -> Path:
    """defines / out: build an A/B variant (e.g. -DCRT_SHADE_MIN_BLOCKS=6) next to the default library;
    select it at run time with CADRAYS_B200_LIB=<path>."""
    if out is None and not force and not _stale():
        return LIB_PATH
    out = out or LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *defines, "-o", str(out), *map(str, _sources()), "-lgomp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    p = build_library(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs,
                      out=Path(outs[0]).resolve() if outs else None)
    print("built", p)
    print("built", build_probe(force="--force" in sys.argv))
