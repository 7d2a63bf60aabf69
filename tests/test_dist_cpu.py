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
