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
