"""GPU-side construction of a bowtie2 index (.bt2 layout) for synthetic workloads.

The reference's `bowtie2-build` (blockwise suffix sorting on the CPU) is
out of scope for the hot path, but the benchmark needs a 3 Gbp index on a box that has no
index files and a few minutes of budget.  This module builds the same arrays on the GPU with
plain torch tensor ops (plumbing, not a hot path): a prefix-doubling suffix sort, then the
BWT / Occ sides, ftab/eftab, SA sample, fchr and zOff exactly as `Ebwt::buildToDisk` lays them
out (bt2_idx.h:2829-3173), plus the `.3/.4` packed reference (reference.cpp:96-200).  On small
genomes the written files are byte-identical to `bowtie2-build-s` output
(tests/test_index_build.py), which is what pins this against the reference.

Suffix order convention (checked against the reference's lambda index): a suffix that is a
proper prefix of another sorts AFTER it, i.e. the end-of-text sentinel is the largest symbol,
and the empty suffix occupies the last BW row.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

OFF_DT = {4: (torch.int64, np.uint32), 8: (torch.int64, np.uint64)}

_T0 = [None]


def _tick(msg, dev=None):
    """phase timing on stderr when BT2G_VERBOSE=1 (synchronises the device)"""
    if not os.environ.get("BT2G_VERBOSE"):
        return
    import sys, time
    if dev is not None and dev.type == "cuda":
        torch.cuda.synchronize(dev)
    now = time.time()
    if _T0[0] is None:
        _T0[0] = now
    print(f"[index_build +{now - _T0[0]:7.1f}s] {msg}", file=sys.stderr, flush=True)


def _group_rank(start: torch.Tensor, rows_base: torch.Tensor = None) -> torch.Tensor:
    """For a bool vector marking group starts (start[0] is True): for every position the
    position of its group's first element.  cumsum + gather (torch.cummax over a single
    multi-billion-element row runs in one thread block)."""
    gid = torch.cumsum(start, 0, dtype=torch.int64) - 1
    firsts = torch.nonzero(start).flatten()
    return firsts[gid]


def _records(contigs: List[torch.Tensor]):
    """RefRecord list (ref_read.h:60-100): per unambiguous stretch (off = #Ns preceding it
    since the previous stretch / contig start, len, first)."""
    recs = []
    for c in contigs:
        good = (c < 4)
        n = good.numel()
        if n == 0:
            continue
        g = good.to(torch.int8)
        d = torch.diff(g, prepend=g.new_zeros(1), append=g.new_zeros(1))
        starts = torch.nonzero(d == 1).flatten().tolist()
        ends = torch.nonzero(d == -1).flatten().tolist()
        prev_end = 0
        first = True
        for s, e in zip(starts, ends):
            recs.append((s - prev_end, e - s, first))
            prev_end = e
            first = False
        if first:
            recs.append((n, 0, True))
        elif prev_end < n:
            recs.append((n - prev_end, 0, False))   # trailing Ns: an empty record carries their count
    return recs


def _pack_k(c: torch.Tensor, nchars: int, bits: int) -> torch.Tensor:
    """key[i] = chars i..i+nchars-1 packed MSB-first with `bits` bits per char (c is padded)."""
    n_out = c.numel() - 64
    cur = c.to(torch.int64)          # 1-char keys
    have = 1
    # doubling: k_{2m}[i] = k_m[i] << (m*bits) | k_m[i+m]
    while have * 2 <= nchars:
        cur = (cur[: cur.numel() - have] << (have * bits)) | cur[have:]
        have *= 2
    if have < nchars:
        rest = nchars - have
        # append the top `rest` chars of the key starting at i+have
        tail = cur[have:] >> ((have - rest) * bits)
        cur = (cur[: tail.numel()] << (rest * bits)) | tail
    return cur[:n_out].contiguous()


def suffix_array(s: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Suffix array (int64, length n+1, empty suffix last) and its inverse, for codes 0..3.

    21-character 3-bit keys (pad symbol 4 > T encodes "end of text is largest") are bucketed
    by their first two characters and radix-sorted per bucket; groups of equal keys are then
    refined by prefix doubling restricted to the still-tied suffixes."""
    dev = s.device
    n = s.numel()
    K, BITS = 21, 3
    c = torch.cat([s, torch.full((64 + 1,), 4, dtype=torch.uint8, device=dev)])   # position n = empty suffix
    key = _pack_k(c, K, BITS)                                  # n+1 keys
    del c
    _tick("keys packed", dev)
    bucket = (key >> (BITS * (K - 2))).to(torch.uint8)
    sa = torch.empty(n + 1, dtype=torch.int64, device=dev)
    start = torch.zeros(n + 2, dtype=torch.bool, device=dev)  # start[row] = row begins a key group
    start[n + 1] = True
    ofs = 0
    for b in [(c1 << 3) | c2 for c1 in range(5) for c2 in range(5)]:
        idx = torch.nonzero(bucket == b).flatten()
        m = idx.numel()
        if m == 0:
            continue
        kb, perm = torch.sort(key[idx])
        sa[ofs:ofs + m] = idx[perm]
        st = torch.ones(m, dtype=torch.bool, device=dev)
        st[1:] = kb[1:] != kb[:-1]
        start[ofs:ofs + m] = st
        ofs += m
        del idx, kb, perm, st
    del key, bucket
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    # rank of a row = first row of its group
    _tick("bucket sorts done", dev)
    rank_of_row = _group_rank(start[: n + 1])
    _tick("group ranks", dev)
    isa = torch.empty(n + 1, dtype=torch.int64, device=dev)
    isa[sa] = rank_of_row
    single = start[: n + 1] & start[1:]
    act = torch.nonzero(~single).flatten()                     # active rows (ascending)
    grp = rank_of_row[act]
    del rank_of_row, single, start
    _tick(f"isa scattered, {act.numel()} tied suffixes", dev)
    h = K
    while act.numel() > 0:
        pos = sa[act]
        nxt = torch.clamp(pos + h, max=n)
        k2 = isa[nxt]
        gstart = torch.ones_like(grp, dtype=torch.bool)
        gstart[1:] = grp[1:] != grp[:-1]
        gd = torch.cumsum(gstart.to(torch.int64), 0) - 1      # dense group id
        shift = max(32, int(n + 1).bit_length())               # ranks need more than 32 bits on a genome beyond 4 Gbp (.bt2l)
        if int(gd[-1]) >> (63 - shift):
            raise ValueError("suffix sort: too many tied suffix groups for a 63-bit composite key")
        comp = (gd << shift) | k2
        comp_s, perm = torch.sort(comp)
        pos_s = pos[perm]
        sa[act] = pos_s
        ns = torch.ones_like(gstart)
        ns[1:] = comp_s[1:] != comp_s[:-1]
        first_idx = _group_rank(ns)
        new_rank = act[first_idx]
        isa[pos_s] = new_rank
        nxt_start = torch.ones_like(ns)
        nxt_start[:-1] = ns[1:]
        keep = ~(ns & nxt_start)
        act = act[keep]
        grp = new_rank[keep]
        h *= 2
        _tick(f"refined to h={h}, {act.numel()} still tied", dev)
    return sa, isa


@dataclass
class EbwtArrays:
    ebwt: torch.Tensor          # uint8 [numSides*sideSz]
    z_off: int
    fchr: List[int]
    ftab: torch.Tensor          # int64 [ftabLen] (stored width applied at write/upload)
    eftab: torch.Tensor         # int64 [2*ftabChars]
    offs: Optional[torch.Tensor]  # int64 [offsLen]


def _build_ebwt(s: torch.Tensor, off_size: int, off_rate: int, ftab_chars: int, want_offs: bool) -> EbwtArrays:
    dev = s.device
    n = s.numel()
    _tick(f"suffix array of {n} symbols", dev)
    sa, isa = suffix_array(s)
    z_off = int(isa[0])
    _tick("suffix array done", dev)
    # everything that needs sa / isa first, so both can be released before the k-mer pass
    # (each is 8 bytes per base: 24 GB at 3 Gbp)
    short_pos = torch.arange(max(n - ftab_chars + 1, 0), n + 1, device=dev)
    short_rows = sorted(isa[short_pos].tolist())
    del isa
    short_set = set(short_rows)
    absorb_at = []                                   # (text position of the next long suffix | None, run length)
    run = 0
    for r in short_rows:
        run += 1
        nxt = r + 1
        if nxt in short_set:
            continue
        absorb_at.append((None if nxt > n else int(sa[nxt]), run))
        run = 0
    offs = sa[:: (1 << off_rate)].clone() if want_offs else None
    # BWT: char preceding each suffix; "$" (SA == 0) stored as A and not counted (bt2_idx.h:2958-2971)
    sa -= 1
    sa.clamp_(min=0)
    bwt = s[sa]
    del sa
    bwt[z_off] = 0
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    side_sz = 16 * off_size
    side_bwt_sz = side_sz - 4 * off_size
    side_bwt_len = side_bwt_sz * 4
    bwt_sz = n // 4 + 1
    num_sides = (bwt_sz + side_bwt_sz - 1) // side_bwt_sz
    padded = torch.zeros(num_sides * side_bwt_len, dtype=torch.uint8, device=dev)
    padded[: n + 1] = bwt
    del bwt
    v = padded.view(num_sides, side_bwt_len)
    # occurrence counts before each side (padding rows count as A; the "$" row does not)
    cnt = torch.stack([(v == ch).sum(dim=1) for ch in range(4)], dim=1).to(torch.int64)   # [numSides, 4]
    cnt[z_off // side_bwt_len, 0] -= 1
    occ = torch.cumsum(cnt, 0) - cnt
    q = padded.view(num_sides, side_bwt_sz, 4)
    packed = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).contiguous()
    del padded, v, q
    if off_size == 4:
        occ_b = occ.to(torch.int32).contiguous().view(torch.uint8).view(num_sides, 16)
    else:
        occ_b = occ.contiguous().view(torch.uint8).view(num_sides, 32)
    ebwt = torch.cat([packed, occ_b], dim=1).contiguous().view(-1)
    _tick("sides packed", dev)
    del packed, occ_b, occ, cnt
    # fchr (bt2_idx.h:3089-3105)
    cc = [int((s == ch).sum()) for ch in range(4)]
    fchr = [0, cc[0], cc[0] + cc[1], cc[0] + cc[1] + cc[2], n]
    # ftab / eftab (bt2_idx.h:2973-3006, :3107-3160)
    ftab_len = (1 << (2 * ftab_chars)) + 1

    def kmer_at(p):
        v_ = 0
        for c_ in s[p:p + ftab_chars].tolist():
            v_ = (v_ << 2) | c_
        return v_

    absorb = np.zeros(ftab_len, dtype=np.int64)
    for p_, run_ in absorb_at:
        absorb[ftab_len - 1 if p_ is None else kmer_at(p_)] = run_
    cpad = torch.cat([s, torch.zeros(64, dtype=torch.uint8, device=dev)])
    kmer = _pack_k(cpad, ftab_chars, 2)[: max(n - ftab_chars + 1, 0)]
    del cpad
    cnt10 = torch.bincount(kmer, minlength=ftab_len - 1).to(torch.int64) if kmer.numel() else \
        torch.zeros(ftab_len - 1, dtype=torch.int64, device=dev)
    del kmer
    cnt_np = cnt10.cpu().numpy()
    csum = np.concatenate([[0], np.cumsum(cnt_np)])            # sum_{k<i} cnt[k]
    hi = csum + np.cumsum(absorb)
    lo = hi - absorb
    mask = (1 << (8 * off_size)) - 1
    ftab = lo.astype(np.uint64)
    eftab = np.zeros(2 * ftab_chars, dtype=np.uint64)
    e = 0
    for i in np.nonzero(absorb)[0]:
        if i == 0:
            continue
        eftab[2 * e] = lo[i]; eftab[2 * e + 1] = hi[i]
        ftab[i] = (e ^ mask) & mask
        e += 1
    _tick("ftab done", dev)
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return EbwtArrays(ebwt, z_off, fchr, torch.from_numpy(ftab.astype(np.int64)), torch.from_numpy(eftab.astype(np.int64)), offs)


@dataclass
class BuiltIndex:
    off_size: int
    line_rate: int
    off_rate: int
    ftab_chars: int
    len: int
    names: List[str]
    plen: List[int]
    rstarts: List[int]
    recs: list
    fw: EbwtArrays
    bw: EbwtArrays
    ref_buf: torch.Tensor       # uint8 2-bit packed joined text
    tensors: dict = field(default_factory=dict)

    # ---- hand the arrays to libbt2g (they already live on the GPU) ------------------------
    def device_desc(self, device):
        """dict for Bt2Gpu.load_index_device(); tensors are kept alive in self.tensors."""
        np_dt = np.uint32 if self.off_size == 4 else np.uint64
        t_dt = torch.int32 if self.off_size == 4 else torch.int64

        def off_t(x):
            if isinstance(x, torch.Tensor):
                return x.to(device=device).to(t_dt).contiguous()
            return torch.from_numpy(np.asarray(x, dtype=np.uint64).astype(np_dt).view(np.int32 if self.off_size == 4 else np.int64)).to(device)

        T = self.tensors
        T["plen"] = off_t(self.plen); T["rstarts"] = off_t(self.rstarts)
        T["ebwt_fw"] = self.fw.ebwt.to(device); T["ebwt_bw"] = self.bw.ebwt.to(device)
        T["ftab_fw"] = off_t(self.fw.ftab); T["eftab_fw"] = off_t(self.fw.eftab)
        T["ftab_bw"] = off_t(self.bw.ftab); T["eftab_bw"] = off_t(self.bw.eftab)
        T["offs"] = off_t(self.fw.offs)
        T["rec_off"] = off_t([r[0] for r in self.recs]); T["rec_len"] = off_t([r[1] for r in self.recs])
        T["rec_first"] = torch.tensor([1 if r[2] else 0 for r in self.recs], dtype=torch.uint8, device=device)
        T["ref_buf"] = self.ref_buf.to(device)
        d = dict(off_size=self.off_size, line_rate=self.line_rate, off_rate=self.off_rate, ftab_chars=self.ftab_chars,
                 len=self.len, n_pat=len(self.plen), n_frag=len(self.rstarts) // 3,
                 z_off_fw=self.fw.z_off, z_off_bw=self.bw.z_off, fchr=self.fw.fchr, n_recs=len(self.recs))
        for k in ("plen", "rstarts", "ebwt_fw", "ebwt_bw", "ftab_fw", "eftab_fw", "ftab_bw", "eftab_bw", "offs",
                  "rec_off", "rec_len", "rec_first", "ref_buf"):
            d[k] = T[k].data_ptr()
        return d

    # ---- write <base>.{1,2,3,4,rev.1,rev.2}.bt2[l] (bt2_io.cpp:700-930 writeFromMemory layout) ----
    def write_files(self, base: str):
        ext = "bt2" if self.off_size == 4 else "bt2l"
        o = "<I" if self.off_size == 4 else "<Q"
        np_dt = np.uint32 if self.off_size == 4 else np.uint64

        def offs_bytes(x):
            if isinstance(x, torch.Tensor):
                x = x.cpu().numpy()
            return np.asarray(x).astype(np.uint64).astype(np_dt).tobytes()

        def write1(path, e: EbwtArrays, flags: int, rstarts):
            with open(path, "wb") as f:
                f.write(struct.pack("<i", 1))
                f.write(struct.pack(o, self.len))
                f.write(struct.pack("<iiiii", self.line_rate, 2, self.off_rate, self.ftab_chars, flags))
                f.write(struct.pack(o, len(self.plen)))
                f.write(offs_bytes(self.plen))
                f.write(struct.pack(o, len(rstarts) // 3))
                f.write(offs_bytes(rstarts))
                eb = e.ebwt.cpu().numpy()
                eb.tofile(f)
                f.write(struct.pack(o, e.z_off))
                f.write(offs_bytes(e.fchr))
                f.write(offs_bytes(e.ftab))
                f.write(offs_bytes(e.eftab))
                for i, nm in enumerate(self.names):
                    f.write(nm.encode() + b"\n")
                f.write(b"\0")

        write1(f"{base}.1.{ext}", self.fw, -1, self.rstarts)
        write1(f"{base}.rev.1.{ext}", self.bw, -5, self.rstarts_rev())
        with open(f"{base}.2.{ext}", "wb") as f:
            f.write(struct.pack("<i", 1))
            f.write(offs_bytes(self.fw.offs))
        with open(f"{base}.rev.2.{ext}", "wb") as f:
            f.write(struct.pack("<i", 1))
            if self.bw.offs is not None:
                f.write(offs_bytes(self.bw.offs))
        with open(f"{base}.3.{ext}", "wb") as f:
            f.write(struct.pack("<i", 1))
            f.write(struct.pack(o, len(self.recs)))
            for off, ln, first in self.recs:
                f.write(struct.pack(o, off)); f.write(struct.pack(o, ln)); f.write(b"\1" if first else b"\0")
        with open(f"{base}.4.{ext}", "wb") as f:
            self.ref_buf.cpu().numpy().tofile(f)

    def rstarts_rev(self):
        """rstarts of the mirror index: records reversed (Ebwt::szsToDisk with REF_READ_REVERSE,
        bt2_io.cpp:933-959 on reverseRefRecords output).  Never loaded by the aligner
        (bt2_search.cpp:4845-4853); written for file-format completeness."""
        out = []
        tot = 0
        npat = len(self.plen)
        # reversed record list: references in reverse order, stretches in reverse order
        per_ref = []
        cur = None
        for off, ln, first in self.recs:
            if first:
                cur = []
                per_ref.append(cur)
            cur.append((off, ln))
        for ri in range(len(per_ref) - 1, -1, -1):
            stretches = per_ref[ri]
            plen = self.plen[ri]
            # positions of stretches in forward coordinates
            pos, fwd = 0, []
            for off, ln in stretches:
                pos += off
                fwd.append((pos, ln))
                pos += ln
            for p, ln in reversed(fwd):
                if ln == 0:
                    continue
                out += [tot, ri, p]
                tot += ln
        return out


def build_index(contigs: List[torch.Tensor], names: Optional[List[str]] = None, off_size: int = 4,
                off_rate: int = 4, ftab_chars: int = 10, mirror_offs: bool = False) -> BuiltIndex:
    """contigs: uint8 code tensors (0..3, 4 = N) on the device to build on."""
    dev = contigs[0].device
    names = names or [f"chr{i + 1}" for i in range(len(contigs))]
    # joined text = all unambiguous stretches, in order; plen/rstarts as Ebwt::joinToDisk /
    # szsToDisk write them (bt2_idx.h:2730-2745, bt2_io.cpp:933-959)
    recs, parts, plen, rstarts = [], [], [], []
    tot = 0
    for ci, c in enumerate(contigs):
        crecs = _records([c])
        recs += crecs
        pos = 0
        plen.append(0)
        for off, ln, fst in crecs:
            pos += off
            if ln > 0:
                parts.append(c[pos:pos + ln])
                rstarts += [tot, ci, pos]
                tot += ln
            pos += ln
            plen[ci] += off + ln
    s = torch.cat(parts) if len(parts) > 1 else parts[0].clone()
    n = s.numel()
    assert n == tot
    if off_size == 4 and n >= (1 << 32) - 200:
        raise ValueError("text too long for a small (.bt2) index")
    fw = _build_ebwt(s, off_size, off_rate, ftab_chars, True)
    bw = _build_ebwt(torch.flip(s, [0]), off_size, off_rate, ftab_chars, mirror_offs)
    # .4: 2-bit packed joined text (reference.cpp:170-260)
    pad = (-n) % 4
    sp = torch.cat([s, torch.zeros(pad, dtype=torch.uint8, device=dev)]) if pad else s
    q = sp.view(-1, 4)
    ref_buf = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).contiguous()
    return BuiltIndex(off_size, 6 if off_size == 4 else 7, off_rate, ftab_chars, n, names, plen, rstarts, recs, fw, bw, ref_buf)
