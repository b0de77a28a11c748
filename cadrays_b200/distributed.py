"""Multi-GPU plumbing: one process per GPU, scene replicated, disjoint sample ranges,
float accumulation buffers summed with one all-reduce.

The reference has no distribution at all (single process, single GL context,
src/Launcher/AppViewer.cxx:593); this is new work.  torch.distributed is only the
plumbing: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Tuple


def sample_range(rank: int, world: int, total_samples: int) -> Tuple[int, int]:
    """[first, first+count) of the per-pixel sample indices rank `rank` renders.

    Contiguous blocks, remainder spread over the low ranks; the union over ranks is
    exactly [0, total_samples) and the stream of sample s of pixel p depends only on
    (frame_seed0, s, p), so N ranks reproduce the 1-GPU sample set."""
    if not (0 <= rank < world) or total_samples < 0:
        raise ValueError("bad rank/world/total")
    base, rem = divmod(total_samples, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def step_sample_start(step: int, rank: int, world: int, spp_per_step: int) -> int:
    """Weak-scaling schedule used by bench.py: step s covers samples
    [s*world*spp, (s+1)*world*spp); rank r owns the r-th block of spp."""
    return (step * world + rank) * spp_per_step


def adaptive_tile_share(count: int, k: int, rank: int, world: int):
    """Adaptive screen sampling over several GPUs (csrc/group.inl, k_adaptive_allocate): a tile whose pixels hold
    `count` samples receives `k` more -- global sample indices count .. count + k - 1 -- and member `rank` of `world`
    renders the ones congruent to its rank.  Returns (first local offset t0, number of own samples): the member's
    samples are count + t0 + i * world for i < number.  The same integers the kernels compute."""
    if not (0 <= rank < world) or count < 0 or k < 0:
        raise ValueError("bad rank/world/count")
    below = lambda x: (x + world - 1 - rank) // world          # sample indices < x congruent to rank
    return (rank + world - count % world) % world, below(count + k) - below(count)


def init_from_env(backend: str | None = None):
    """torch.distributed bootstrap from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_sum_(tensor):
    """In-place sum over ranks of an accumulation tensor (rgb sums + sample counts)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
