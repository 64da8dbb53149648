"""World-size-2 `gloo` test of the multi-GPU host logic (approx_counter_b200/sharded.py):
reads are sharded per rank, every rank counts its shard for ALL k-mers, one
all-reduce sums the count vectors.  On CPU the per-shard counts come from the
oracle (the checker standing in for the scan kernel); what is under test is the
shard arithmetic and the collective."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align():
    from approx_counter_b200.sharded import shard_bounds
    for n in (0, 1, 31, 32, 33, 1000, 100000, 1000003):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert lo == prev and lo <= hi <= n
                assert lo % 32 == 0 or lo == n
                prev = hi
            assert prev == n
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(s for s in sizes if s or True) <= max(32 * world, max(sizes))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from approx_counter_b200.sharded import allreduce_counts, shard_bounds
    from oracle import orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)  # same data on every rank
        n, L, k = 333, 60, 12
        sample = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, L))
        needle = sample[0, 10:22].copy()
        sample[::3, 20:32] = needle
        kmers = np.array([orc.dna2int(needle.tobytes().decode())] +
                         [int(x) & ((1 << (2 * k)) - 1) for x in rng.integers(0, 1 << 62, 7)], np.uint64)
        lo, hi = shard_bounds(n, rank, world)
        codes, offs = orc.encode_matrix(sample[lo:hi])
        local = orc.error_count(codes, offs, kmers, k, fast=True)
        t = torch.from_numpy(local.view(np.int64).copy())
        allreduce_counts(t)
        codes, offs = orc.encode_matrix(sample)
        want = orc.error_count(codes, offs, kmers, k, fast=True)
        ok = np.array_equal(t.numpy().view(np.uint64), want) and int(want[0]) >= 3 * len(range(0, n, 3))
        out.put((rank, bool(ok), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_sharded_counts_allreduce_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2][1] == res[1][2][0]  # contiguous shards
