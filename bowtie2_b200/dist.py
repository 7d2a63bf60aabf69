"""Multi-GPU plumbing: reads shard across ranks with no per-step collective;
the index is loaded/built by ONE rank and every array is broadcast once at start-up
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

SCALARS = ("off_size", "line_rate", "off_rate", "ftab_chars", "len", "n_pat", "n_frag", "z_off_fw", "z_off_bw", "n_recs")
ARRAYS = ("plen", "rstarts", "ebwt_fw", "ebwt_bw", "ftab_fw", "eftab_fw", "ftab_bw", "eftab_bw", "offs",
          "rec_off", "rec_len", "rec_first", "ref_buf")


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of read ids owned by `rank` (the reference's -s/--skip, -u/--upto manual
    sharding, bt2_search.cpp:3278, made automatic).  Blocks differ in size by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_index(built, src: int, device: torch.device):
    """Rank `src` passes a BuiltIndex (bowtie2_b200.index_build); every rank returns
    (desc, tensors): the bt2g_index_host description with device pointers, and the tensors that
    back it (keep them alive).  One broadcast per array, no other communication."""
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        desc = built.device_desc(device)
        meta[0] = {k: desc[k] for k in SCALARS}
        meta[0]["fchr"] = list(desc["fchr"])
        meta[0]["shapes"] = {k: (tuple(built.tensors[k].shape), str(built.tensors[k].dtype).split(".")[1]) for k in ARRAYS}
    dist.broadcast_object_list(meta, src=src)
    m = meta[0]
    tensors = {}
    for k in ARRAYS:
        shape, dt = m["shapes"][k]
        t = built.tensors[k] if rank == src else torch.empty(shape, dtype=getattr(torch, dt), device=device)
        dist.broadcast(t, src=src)
        tensors[k] = t
    desc = {k: m[k] for k in SCALARS}
    desc["fchr"] = m["fchr"]
    for k in ARRAYS:
        desc[k] = tensors[k].data_ptr()
    return desc, tensors


def deal_blocks(n_blocks: int, block_spec, result_spec, get_block, align, put_result, device: torch.device, depth: int = 2, src: int = 0,
                sync=None):
    """ONE reader deals fixed-size blocks of reads round-robin to the ranks, ONE ordered writer collects the
    results.  Rank `src` is both: it calls get_block(k) -> tuple of tensors (shapes / dtypes = block_spec) for every block, aligns
    its own share (k % world == src) and hands every block's results to put_result(k, tensors) -- in block order per rank, the
    caller's sink reorders across ranks (blocks of different ranks finish in any order).  The other ranks receive their blocks
    (k % world == rank), run align(tensors) -> tuple of tensors (result_spec) and send the results back.  Point-to-point only
    (NCCL send / recv over NVLink on GPUs, gloo in the CPU test): `depth` blocks are in flight per peer, so a rank's next block
    arrives while it aligns the current one.  *_spec: list of (shape, torch dtype).  sync (optional): called after a received block
    is complete on the communication stream and before align() -- a host-side wait when align() works on streams torch does not
    order against its own (the engines' streams).  Returns the number of blocks this rank aligned."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = [k for k in range(n_blocks) if k % world == rank]
    if world == 1:
        for k in mine:
            put_result(k, align(get_block(k)))
        return len(mine)

    def wait_all(hs):
        for h in hs:
            h.wait()

    if rank != src:
        bufs = [[torch.empty(s, dtype=d, device=device) for s, d in block_spec] for _ in range(depth)]
        handles = [None] * len(mine)
        for i in range(min(depth, len(mine))):
            handles[i] = [dist.irecv(t, src=src) for t in bufs[i % depth]]
        for i, k in enumerate(mine):
            wait_all(handles[i])
            if sync:
                sync()
            res = align(tuple(bufs[i % depth]))
            wait_all([dist.isend(t.contiguous(), dst=src) for t in res])
            if i + depth < len(mine):                       # (its buffer is free again)
                handles[i + depth] = [dist.irecv(t, src=src) for t in bufs[i % depth]]
        return len(mine)

    # the dealer / collector
    peers = [r for r in range(world) if r != src]
    theirs = {r: [k for k in range(n_blocks) if k % world == r] for r in peers}
    sent = {r: 0 for r in peers}
    got = {r: 0 for r in peers}
    inflight = {r: [] for r in peers}                        # send handles + the tensors they read (kept alive)
    rbuf = {r: [torch.empty(s, dtype=d, device=device) for s, d in result_spec] for r in peers}

    def feed(r):
        while sent[r] < len(theirs[r]) and sent[r] < got[r] + depth:
            blk = tuple(t.contiguous() for t in get_block(theirs[r][sent[r]]))
            inflight[r].append(([dist.isend(t, dst=r) for t in blk], blk))
            sent[r] += 1

    rounds = max([len(mine)] + [len(v) for v in theirs.values()])
    for i in range(rounds):
        for r in peers:
            feed(r)
        if i < len(mine):
            put_result(mine[i], align(get_block(mine[i])))
        for r in peers:
            if got[r] < len(theirs[r]):
                wait_all([dist.irecv(t, src=r) for t in rbuf[r]])
                put_result(theirs[r][got[r]], tuple(rbuf[r]))
                got[r] += 1
                hs, _ = inflight[r].pop(0)
                wait_all(hs)
    for r in peers:                                          # (rounds covers every block; nothing is left)
        assert got[r] == len(theirs[r]) and not inflight[r]
    return len(mine)
