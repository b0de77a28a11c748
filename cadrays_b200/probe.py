"""Memory-hierarchy probe (libcrt_probe.so, csrc/probe/mem_probe.cu): measured L2 / HBM read bandwidth for
sequential vectors and for scattered 64-byte records (the shape of a BVH node fetch), on the device the bench
runs on.  SURVEY 8(d) asks for the L2 figures as the memory roofline of the L2-resident configs (C1-C4)."""
from __future__ import annotations

import ctypes as C

from .build import PROBE_PATH

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not PROBE_PATH.exists():
            raise RuntimeError(f"{PROBE_PATH} is missing: build it with `python -m cadrays_b200.build`")
        L = C.CDLL(str(PROBE_PATH))
        L.crt_probe_bandwidth.restype = C.c_double
        L.crt_probe_bandwidth.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_int]
        L.crt_probe_device.restype = C.c_int
        L.crt_probe_device.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def device_info(device: int = 0) -> dict:
    l2, sm, khz = C.c_size_t(), C.c_int(), C.c_int()
    if _load().crt_probe_device(device, C.byref(l2), C.byref(sm), C.byref(khz)) != 0:
        raise RuntimeError("crt_probe_device failed")
    return {"l2_bytes": l2.value, "sm_count": sm.value, "sm_clock_mhz": khz.value / 1000.0}


def bandwidth(device: int, nbytes: int, mode: str, reps: int = 3) -> float:
    """GB/s of `mode` in {"sequential", "records", "records_dependent"} over a working set of nbytes."""
    m = {"sequential": 0, "records": 1, "records_dependent": 2}[mode]
    v = _load().crt_probe_bandwidth(device, int(nbytes), m, reps)
    if v < 0:
        raise RuntimeError(f"crt_probe_bandwidth failed ({v})")
    return v


def sweep(device: int = 0, sizes_mb=(16, 32, 64, 96, 120, 160, 256, 1024, 4096)) -> dict:
    """Bandwidth against working-set size: the knee is the usable L2 capacity."""
    out = {"device": device_info(device), "unit": "GB/s", "sets": []}
    for mb in sizes_mb:
        row = {"working_set_mb": mb}
        for mode in ("sequential", "records", "records_dependent"):
            row[mode] = bandwidth(device, mb << 20, mode)
        out["sets"].append(row)
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(sweep(), indent=1))
