"""world_size-2 gloo test of the N>1 path on CPU: disjoint sample ranges per rank + one all-reduce
of the float accumulation buffers reproduce the single-process image.  The renderer
stand-in on CPU is the oracle (no GPU here); the partition / reduce logic under test is the product's
cadrays_b200.distributed."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from cadrays_b200 import distributed as D

REPO = Path(__file__).resolve().parent.parent


def test_sample_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 64, 4096, 4099):
            got = []
            for r in range(world):
                first, n = D.sample_range(r, world, total)
                got.extend(range(first, first + n))
            assert got == list(range(total))
            sizes = [D.sample_range(r, world, total)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.sample_range(2, 2, 10)
    # bench schedule: blocks of spp, disjoint across ranks and steps
    seen = set()
    for step in range(3):
        for r in range(4):
            s0 = D.step_sample_start(step, r, 4, 5)
            blk = set(range(s0, s0 + 5))
            assert not (blk & seen)
            seen |= blk
    assert seen == set(range(60))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    from cadrays_b200 import scenes
    from cadrays_b200.view import V3d_View
    from oracle.oracle_ffi import OracleScene
    r, _, w = D.init_from_env("gloo")
    desc = scenes.cornell_box(40, 32, depth=4, sphere_res=(12, 6))
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    o = OracleScene(v.ExportBVH())
    o.configure(desc)
    total = 7
    first, n = D.sample_range(r, w, total)
    acc = o.render(40, 32, n, first_sample=first)
    t = torch.from_numpy(acc)
    D.allreduce_sum_(t)
    if r == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process(tmp_path, product_lib, oracle_lib):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    from cadrays_b200 import scenes
    from cadrays_b200.view import V3d_View
    from oracle.oracle_ffi import OracleScene
    desc = scenes.cornell_box(40, 32, depth=4, sphere_res=(12, 6))
    v = V3d_View(host_only=True)
    desc.apply(v, with_target=False)
    o = OracleScene(v.ExportBVH())
    o.configure(desc)
    single = o.render(40, 32, 7)
    assert np.array_equal(reduced[..., 3], single[..., 3])            # every pixel got all 7 samples
    assert np.allclose(reduced, single, rtol=2e-6, atol=1e-6)         # float summation order only


def test_sample_partition_properties_randomised():
    """sample_range / step_sample_start over random worlds and totals (hypothesis): the ranges tile [0, total)
    exactly, in rank order, sizes differ by at most one; the weak-scaling schedule never hands a sample index to two
    (step, rank) pairs."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(world=st.integers(1, 64), total=st.integers(0, 100_000))
    def ranges(world, total):
        spans = [D.sample_range(r, world, total) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f1 == f0 + c0
        counts = [c for _, c in spans]
        assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)

    @settings(max_examples=100, deadline=None)
    @given(world=st.integers(1, 16), spp=st.integers(1, 64), steps=st.integers(1, 20))
    def schedule(world, spp, steps):
        seen = set()
        for s in range(steps):
            for r in range(world):
                first = D.step_sample_start(s, r, world, spp)
                block = range(first, first + spp)
                assert seen.isdisjoint(block)
                seen.update(block)
        assert seen == set(range(steps * world * spp))

    ranges()
    schedule()


def test_adaptive_tile_share_partitions_every_tile():
    """The dealing rule of adaptive sampling over a group: for any tile state the members' shares are disjoint, cover
    the tile's new samples exactly, differ in size by at most one, and follow on from wave to wave."""
    import random
    rnd = random.Random(3)
    for _ in range(3000):
        world = rnd.randint(1, 9)
        count, k = rnd.randint(0, 60), rnd.randint(0, 40)
        seen = []
        sizes = []
        for r in range(world):
            t0, n = D.adaptive_tile_share(count, k, r, world)
            own = [count + t0 + i * world for i in range(n)]
            assert all(count <= g < count + k and g % world == r for g in own)
            seen += own
            sizes.append(n)
        assert sorted(seen) == list(range(count, count + k))
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.adaptive_tile_share(0, 1, 2, 2)
