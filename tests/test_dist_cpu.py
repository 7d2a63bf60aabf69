"""CPU, world_size 2, gloo: read sharding and the one-shot index broadcast."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bowtie2_b200 import synth
    from bowtie2_b200.dist import ARRAYS, broadcast_index, shard_range
    from bowtie2_b200.index_build import build_index
    built = None
    if rank == 0:
        g = synth.make_genome(2, 3000, seed=4, repeat_frac=0.0, n_gap=11)
        built = build_index([torch.from_numpy(c) for c in g])
    desc, tensors = broadcast_index(built, 0, torch.device("cpu"))
    h = hashlib.sha256()
    for k in ARRAYS:
        h.update(tensors[k].numpy().tobytes())
    lo, hi = shard_range(1001, rank, world)
    q.put((rank, h.hexdigest(), {k: desc[k] for k in ("len", "n_pat", "n_frag", "z_off_fw", "z_off_bw", "off_size")}, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    world = 2
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, h0, d0, lo0, hi0), (r1, h1, d1, lo1, hi1) = out
    assert h0 == h1 and d0 == d1 and d0["len"] == 2 * 3000 - 2 * 11
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)


def test_shard_range_partitions():
    from bowtie2_b200.dist import shard_range
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _deal_worker(rank, world, port, q, n_blocks):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bowtie2_b200.dist import deal_blocks
    B = 64
    spec_in = [((B, 10), torch.uint8), ((B,), torch.int64)]
    spec_out = [((B,), torch.int64), ((B, 2), torch.int32)]
    got = {}

    def get_block(k):                       # (rank 0 only) block k: rows filled with k, ids k * B ...
        return torch.full((B, 10), k % 251, dtype=torch.uint8), torch.arange(k * B, (k + 1) * B, dtype=torch.int64)

    def align(t):                           # stand-in for the engine: a function of the block only, tagged with the aligning rank
        rows, ids = t
        return ids * 3 + rows[:, 0].to(torch.int64), torch.stack([torch.full((B,), rank, dtype=torch.int32), rows.sum(1).to(torch.int32)], 1)

    def put(k, res):
        got[k] = (res[0].clone(), res[1].clone())
    n = deal_blocks(n_blocks, spec_in, spec_out, get_block, align, put, torch.device("cpu"), depth=2)
    if rank == 0:
        ok = sorted(got) == list(range(n_blocks))
        for k in range(n_blocks):
            a, b = got[k]
            ok = ok and bool((a == torch.arange(k * B, (k + 1) * B) * 3 + k % 251).all()) and bool((b[:, 0] == k % world).all()) \
                and bool((b[:, 1] == 10 * (k % 251)).all())
        q.put((rank, n, ok))
    else:
        q.put((rank, n, True))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_blocks", [(2, 7), (3, 10), (2, 1)])
def test_one_reader_deals_blocks_one_writer_collects(world, n_blocks):
    """SURVEY 8e topology (bowtie2_b200/dist.py: deal_blocks) on gloo: every block is aligned exactly once, by rank k % world, and its
    results reach the collector on rank 0"""
    port = 31500 + (os.getpid() * 7 + world * 13 + n_blocks) % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_deal_worker, args=(r, world, port, q, n_blocks)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, _, ok in out)
    assert sum(n for _, n, _ in out) == n_blocks and [n for _, n, _ in out] == [len(range(r, n_blocks, world)) for r in range(world)]
