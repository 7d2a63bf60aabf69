"""The Bt2Gpu methods the policy backend uses, answered by the CPU oracle in the array conventions of the real entry
points (those conventions are what the -m gpu parity tests pin).  Lets the CPU suite exercise
bowtie2_b200.policy_backend_gpu (grouping, chunking, conversions) without a device; test infrastructure only."""
import numpy as np

from bowtie2_b200 import policy
from bowtie2_b200.lib import DP_ALN, DP_CAND, DP_SUMMARY, MM_HIT, UNGAPPED_RESULT
from bowtie2_b200.policy_engine import Aln, aln_to_ops
from oracle_lib import extend_both, oracle_dp, oracle_one_mm, oracle_ungapped


class FakeGpu:
    def __init__(self, O, off_size=4):
        self.O = O
        self.local = False
        self._off_size = off_size
        self.calls = 0
        from bowtie2_b200.lib import load_library
        self._lib = load_library()               # host-side formatting / parsing entry points are real

    def load_index_host(self, image):
        pass

    def policy_backend_table(self):
        return backend_table(self)

    def close(self):
        pass

    def set_scoring(self, local=False):
        self.local = local

    def info(self):
        return {"off_size": self._off_size}

    @staticmethod
    def _reads(batch):
        return [(batch.seq[int(batch.off[i]):int(batch.off[i + 1])],
                 None if batch.qual is None else batch.qual[int(batch.off[i]):int(batch.off[i + 1])]) for i in range(batch.n)]

    def exact_sweep(self, batch, nofw=False, norc=False):
        self.calls += 1
        mine = np.zeros((batch.n, 2), dtype=np.uint8)
        ee = np.zeros((batch.n, 4), dtype=np.uint64)
        for i, (c, _) in enumerate(self._reads(batch)):
            _, m, tb = self.O.exact_sweep(c, nofw, norc)
            mine[i], ee[i] = m, tb
        return mine, ee

    def one_mm(self, batch, minsc, strand_mask=3, max_hits=16):
        self.calls += 1
        hits = np.zeros((batch.n, 4, max_hits), dtype=MM_HIT)
        counts = np.zeros((batch.n, 4), dtype=np.int32)
        minsc = np.broadcast_to(minsc, (batch.n,))
        mask = np.broadcast_to(strand_mask, (batch.n,))
        code = {ord(c): i for i, c in enumerate("ACGTN")}
        for i, (c, q) in enumerate(self._reads(batch)):
            for (t, b, p, ch, qch, s, fw) in oracle_one_mm(self.O, self.local, c, q, int(minsc[i]), not (mask[i] & 1), not (mask[i] & 2)):
                task = 0 if fw else 2
                hits[i, task, counts[i, task]] = (t, b, p, code[ch], code[qch], s)
                counts[i, task] += 1
        return hits, counts

    def seed_search(self, batch, seed_len, interval, offset, max_seeds, nofw=False, norc=False):
        self.calls += 1
        interval = np.broadcast_to(interval, (batch.n,))
        offset = np.broadcast_to(offset, (batch.n,))
        out = np.zeros((batch.n, 2, max_seeds, 4), dtype=np.uint64)
        ns = np.zeros(batch.n, dtype=np.int32)
        for i, (c, q) in enumerate(self._reads(batch)):
            n, o = self.O.seed_search(c, seed_len, int(interval[i]), int(offset[i]), max_seeds, nofw, norc, quals=q)
            ns[i] = n
            out[i] = o
        return out, ns

    def extend_exact(self, batch, seed_len, interval, offset, max_seeds, ranges):
        self.calls += 1
        interval = np.broadcast_to(interval, (batch.n,))
        offset = np.broadcast_to(offset, (batch.n,))
        out = np.zeros((batch.n, 2, max_seeds, 2), dtype=np.uint8)
        for i, (c, _) in enumerate(self._reads(batch)):
            n = policy.n_seeds(len(c), min(seed_len, len(c)), int(interval[i]), int(offset[i]))
            for strand in range(2):
                for k in range(min(n, max_seeds)):
                    rg = ranges[i, strand, k]
                    if rg[1] > rg[0]:
                        out[i, strand, k] = extend_both(self.O, c, strand == 0, int(offset[i]) + k * int(interval[i]), min(seed_len, len(c)), rg)
        return out

    def resolve(self, rows, hitlen, reject_straddle=False):
        self.calls += 1
        rows = np.asarray(rows, dtype=np.uint64)
        n = len(rows)
        hitlen = np.broadcast_to(hitlen, (n,))
        joined = self.O.get_offset(rows)
        tidx, textoff, tlen = (np.zeros(n, dtype=np.uint64) for _ in range(3))
        flags = np.zeros(n, dtype=np.uint8)
        for i in range(n):
            ok, ti, to, tl, st = self.O.joined_to_text(int(hitlen[i]), int(joined[i]), int(reject_straddle))
            flags[i] = (1 if st else 0) | (0 if ok else 2)
            if ok:
                tidx[i], textoff[i], tlen[i] = ti, to, tl
        return joined, tidx, textoff, tlen, flags

    def get_stretch(self, tidx, off, count, stride):
        self.calls += 1
        out = np.full((len(tidx), stride), 4, dtype=np.uint8)
        for i in range(len(tidx)):
            out[i, :int(count[i])] = self.O.get_stretch(int(tidx[i]), int(off[i]), int(count[i]))
        return out

    def ungapped(self, batch, probs, want_mask=True):
        self.calls += 1
        rd = self._reads(batch)
        out = np.zeros(len(probs), dtype=UNGAPPED_RESULT)
        stride = int(batch.lengths().max())
        mask = np.zeros((len(probs), stride), dtype=np.uint8)
        for k, p in enumerate(probs):
            c, q = rd[int(p["read_idx"])]
            rc, d = oracle_ungapped(self.O, self.local, c, q, bool(p["fw"]), int(p["tidx"]), int(p["refoff"]), int(p["reflen"]), int(p["ohang"]),
                                    int(p["minsc"]))
            out[k]["status"] = rc
            if rc == 1:
                out[k]["score"], out[k]["rowi"], out[k]["rowf"] = d["score"], d["rowi"], d["rowf"]
                out[k]["ns"], out[k]["refns"], out[k]["nedits"] = d["ns"], d["refns"], d["nedits"]
                mask[k, :len(c)] = d["mask"]
        return out, mask

    def dp_extend(self, batch, probs, max_cands=128, max_alns=4, max_ops=None):
        self.calls += 1
        rd = self._reads(batch)
        n = len(probs)
        max_ops = max_ops or int(batch.lengths().max()) + 64
        summ = np.zeros(n, dtype=DP_SUMMARY)
        cands = np.zeros((n, max_cands), dtype=DP_CAND)
        alns = np.zeros((n, max_alns), dtype=DP_ALN)
        ops = np.zeros((n, max_alns, max_ops), dtype=np.uint8)
        for k, p in enumerate(probs):
            c, q = rd[int(p["read_idx"])]
            fw = bool(p["fw"])
            rect = policy.DPRect(int(p["refl"]), int(p["refr"]), 0, 0, int(p["triml"]), 0, int(p["corel"]), int(p["corer"]), 0)
            d = oracle_dp(self.O, self.local, c, q, fw, int(p["tidx"]), rect, int(p["minsc"]), int(p["nceil"]), max_cands=65536,
                          max_alns=256, max_edits=65536, attempts=True)
            summ[k]["found"], summ[k]["best"], summ[k]["ncand"], summ[k]["naln"] = d["found"], d["best"], d["ncand"], d["naln"]
            if not d["found"]:
                continue
            flags = 0
            if d["ncand"] > max_cands:
                flags |= 2
            if d["naln"] > max_alns:
                flags |= 4
            summ[k]["flags"] = flags
            for ci, (row, col, score) in enumerate(d["cands"][:max_cands]):
                cands[k][ci] = (score, row, col, 1)                      # start-filtered unless an attempt says otherwise
            cand_of = {}
            for (score, ai), ci in zip(d["attempts"], d["attempt_cands"]):
                if ci < max_cands:
                    cands[k][ci]["fate"] = 3 if ai >= 0 else 2
                if ai >= 0:
                    cand_of[ai] = ci
            for ai, al in enumerate(d["alns"][:max_alns]):
                a = Aln(int(p["tidx"]), al["refoff"], fw, al["score"], len(c), [tuple(e) for e in al["edits"]], al["ns"], 0, False,
                        al["trim5"], al["trim3"])
                o = aln_to_ops(a, c)
                tl = a.trim_left
                refns = int((self.O.get_stretch(int(p["tidx"]), al["refoff"], a.ref_extent) > 3).sum())       # AlnRes::refNs, as the kernel reports it
                alns[k][ai] = (cand_of[ai], al["score"], al["ns"], al["gaps"], refns, tl, al["refoff"] - int(p["refl"]), tl,
                               len(c) - a.ext - tl, len(o))
                ops[k, ai, :len(o)] = o
        return summ, cands, alns, ops


# ---- FakeGpu behind the C function-pointer table of bt2g_policy_align (bt2g_policy_backend) ---------------------------------
import ctypes as C

from bowtie2_b200.lib import DP_PROBLEM, UNGAPPED_PROBLEM, ReadBatch, _PolicyBackend, _Reads


def _arr(ptr, dtype, n):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class _SeedPlanS(C.Structure):
    _fields_ = [("seed_len", C.c_int32), ("max_seeds", C.c_int32), ("nofw", C.c_int32), ("norc", C.c_int32), ("interval", C.c_void_p),
                ("offset", C.c_void_p)]


def _batch(ptr):
    r = C.cast(ptr, C.POINTER(_Reads)).contents
    n = int(r.n_reads)
    off = _arr(r.off, np.uint64, n + 1).copy()
    nb = int(off[-1]) if n else 0
    return ReadBatch(_arr(r.seq, np.uint8, nb).copy(), off, _arr(r.qual, np.uint8, nb).copy() if r.qual else None)


def backend_table(fake: "FakeGpu"):
    """-> (_PolicyBackend, keep-alive list): every entry point of the table answered by `fake` in place of the device"""
    F = dict(_PolicyBackend._fields_)

    def exact_sweep(ctx, reads, nofw, norc, mine, ee):
        b = _batch(reads)
        m, e = fake.exact_sweep(b, bool(nofw), bool(norc))
        _arr(mine, np.uint8, 2 * b.n)[:] = m.reshape(-1)
        _arr(ee, np.uint64, 4 * b.n)[:] = e.reshape(-1)
        return 0

    def seed_search(ctx, reads, plan, out, nseeds):
        b = _batch(reads)
        p = C.cast(plan, C.POINTER(_SeedPlanS)).contents
        o, ns = fake.seed_search(b, p.seed_len, _arr(p.interval, np.int32, b.n), _arr(p.offset, np.int32, b.n), p.max_seeds, bool(p.nofw), bool(p.norc))
        _arr(out, np.uint64, b.n * 2 * p.max_seeds * 4)[:] = o.reshape(-1)
        _arr(nseeds, np.int32, b.n)[:] = ns
        return 0

    def one_mm(ctx, reads, minsc, mask, max_hits, hits, counts):
        from bowtie2_b200.lib import MM_HIT
        b = _batch(reads)
        h, c = fake.one_mm(b, _arr(minsc, np.int32, b.n), _arr(mask, np.uint8, b.n), max_hits)
        _arr(hits, MM_HIT, b.n * 4 * max_hits)[:] = h.reshape(-1)
        _arr(counts, np.int32, b.n * 4)[:] = c.reshape(-1)
        return 0

    def extend_exact(ctx, reads, plan, ranges, out):
        b = _batch(reads)
        p = C.cast(plan, C.POINTER(_SeedPlanS)).contents
        rg = _arr(ranges, np.uint64, b.n * 2 * p.max_seeds * 4).reshape(b.n, 2, p.max_seeds, 4)
        o = fake.extend_exact(b, p.seed_len, _arr(p.interval, np.int32, b.n), _arr(p.offset, np.int32, b.n), p.max_seeds, rg)
        _arr(out, np.uint8, b.n * 2 * p.max_seeds * 2)[:] = o.reshape(-1)
        return 0

    def resolve(ctx, rows, hitlen, n, reject, joined, tidx, textoff, tlen, flags):
        j, t, o, l, f = fake.resolve(_arr(rows, np.uint64, n).copy(), _arr(hitlen, np.uint32, n).copy(), bool(reject))
        _arr(joined, np.uint64, n)[:] = j; _arr(tidx, np.uint64, n)[:] = t; _arr(textoff, np.uint64, n)[:] = o
        _arr(tlen, np.uint64, n)[:] = l; _arr(flags, np.uint8, n)[:] = f
        return 0

    def get_stretch(ctx, tidx, off, count, n, stride, out):
        s = fake.get_stretch(_arr(tidx, np.uint64, n), _arr(off, np.int64, n), _arr(count, np.int32, n), stride)
        _arr(out, np.uint8, n * stride)[:] = s.reshape(-1)
        return 0

    def ungapped(ctx, reads, probs, n, out, mask, stride):
        from bowtie2_b200.lib import UNGAPPED_RESULT
        b = _batch(reads)
        o, m = fake.ungapped(b, _arr(probs, UNGAPPED_PROBLEM, n).copy())
        _arr(out, UNGAPPED_RESULT, n)[:] = o
        mm = _arr(mask, np.uint8, n * stride).reshape(n, stride)
        mm[:, :m.shape[1]] = m[:, :stride]
        return 0

    def dp_extend(ctx, reads, probs, n, max_cands, max_alns, max_ops, summ, cands, alns, ops):
        from bowtie2_b200.lib import DP_ALN, DP_CAND, DP_SUMMARY
        b = _batch(reads)
        s, c, a, o = fake.dp_extend(b, _arr(probs, DP_PROBLEM, n).copy(), max_cands, max_alns, max_ops)
        _arr(summ, DP_SUMMARY, n)[:] = s
        _arr(cands, DP_CAND, n * max_cands)[:] = c.reshape(-1)
        _arr(alns, DP_ALN, n * max_alns)[:] = a.reshape(-1)
        _arr(ops, np.uint8, n * max_alns * max_ops)[:] = o.reshape(-1)
        return 0

    be = _PolicyBackend()
    keep = []
    for name, fn in (("exact_sweep", exact_sweep), ("seed_search", seed_search), ("one_mm", one_mm), ("extend_exact", extend_exact),
                     ("resolve", resolve), ("get_stretch", get_stretch), ("ungapped", ungapped), ("dp_extend", dp_extend)):
        cb = F[name](fn)
        keep.append(cb)
        setattr(be, name, cb)
    be.off_size = fake.info()["off_size"]
    return be, keep
