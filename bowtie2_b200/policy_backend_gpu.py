"""The GPU primitives of include/bt2g.h as a backend of policy_engine (one call per primitive and item: exact, slow).

Every method maps to an entry point whose results are pinned bit-exact against the CPU oracle by the -m gpu tests
(test_fm_gpu / test_onemm / test_extend / test_ungapped / test_dp_gpu), in the conventions those tests establish.
The DP's backtrace attempts come from the per-candidate fates bt2g_dp_extend returns (BT2G_CAND_FAILED / _SUCCEEDED).
NOT YET RUN ON HARDWARE (written after the round's GPU minutes were spent): tests/test_zz_fullsize_gpu.py holds the
staged check (engine over this backend == golden SAM)."""
import numpy as np

from .lib import DP_PROBLEM, UNGAPPED_PROBLEM, Bt2Gpu, ReadBatch, ops_to_edits
from .policy_engine import Aln

CAND_FAILED, CAND_SUCCEEDED = 2, 3


class GpuBackend:
    def __init__(self, gpu: Bt2Gpu, local: bool = False):
        self.gpu = gpu
        self.local = local
        gpu.set_scoring(local=local)
        self.off_size = int(gpu.info()["off_size"])
        self._row = None

    @staticmethod
    def _batch(codes, quals=None):
        return ReadBatch.from_list([np.asarray(codes, dtype=np.uint8)], None if quals is None else [np.asarray(quals, dtype=np.uint8)])

    def exact_sweep(self, codes, nofw=False, norc=False):
        mine, ee = self.gpu.exact_sweep(self._batch(codes), nofw, norc)
        tb = [int(x) for x in ee[0]]
        nelt = max(0, tb[1] - tb[0]) + max(0, tb[3] - tb[2])
        return nelt, [int(mine[0][0]), int(mine[0][1])], tb

    def one_mm(self, codes, quals, minsc, nofw, norc):
        mask = (0 if nofw else 1) | (0 if norc else 2)
        hits, counts = self.gpu.one_mm(self._batch(codes, quals), int(minsc), mask, max_hits=64)
        out = []
        for task in range(4):
            for h in hits[0, task, :counts[0, task]]:
                out.append((int(h["top"]), int(h["bot"]), int(h["pos"]), ord("ACGTN"[int(h["chr"])]), ord("ACGTN"[int(h["qchr"])]),
                            int(h["score"]), int(task < 2)))
        return out

    def seed_search(self, codes, quals, seed_len, interval, offset, nofw=False, norc=False):
        from . import policy
        n = max(1, policy.n_seeds(len(codes), seed_len, interval, offset))
        out, ns = self.gpu.seed_search(self._batch(codes, quals), seed_len, interval, offset, n + 2, nofw, norc)
        return out[0][:, :int(ns[0]), :]

    def extend(self, codes, fw, rdoff, seedlen, rng4):
        # one seed at offset rdoff: the entry point takes the seed layout (offset + k * interval) and a range per seed
        ranges = np.zeros((1, 2, 1, 4), dtype=np.uint64)
        ranges[0, 0 if fw else 1, 0] = [int(x) for x in rng4]
        ext = self.gpu.extend_exact(self._batch(codes), seedlen, max(1, len(codes)), rdoff, 1, ranges)
        e = ext[0, 0 if fw else 1, 0]
        return int(e[0]), int(e[1])

    def resolve(self, row):
        self._row = int(row)
        joined, *_ = self.gpu.resolve(np.array([row], dtype=np.uint64), 1, False)
        return int(joined[0])

    def joined_to_text(self, qlen, off, reject):
        # same row as the preceding resolve(): the entry point resolves and converts in one pass
        joined, tidx, textoff, tlen, flags = self.gpu.resolve(np.array([self._row], dtype=np.uint64), int(qlen), bool(reject))
        assert int(joined[0]) == int(off)
        invalid = (int(flags[0]) >> 1) & 1
        return (not invalid), int(tidx[0]), int(textoff[0]), int(tlen[0]), int(flags[0]) & 1

    def count_ref_ns(self, tidx, off, extent):
        s = self.gpu.get_stretch([tidx], [off], [extent], int(extent))
        return int((s[0][:extent] > 3).sum())

    def ungapped(self, codes, quals, fw, tidx, refoff, tlen, minsc):
        probs = np.zeros(1, dtype=UNGAPPED_PROBLEM)
        probs[0] = (0, int(fw), tidx, refoff, tlen, minsc, 0)
        out, mask = self.gpu.ungapped(self._batch(codes, quals), probs)
        rc = int(out[0]["status"])
        if rc != 1:
            return rc, None
        rdlen = len(codes)
        ref = self.gpu.get_stretch([tidx], [refoff], [rdlen], rdlen)[0]
        seq = codes if fw else np.array([4 if c > 3 else 3 - c for c in codes[::-1]], dtype=np.uint8)
        rowi, rowf = int(out[0]["rowi"]), int(out[0]["rowf"])
        ext = rowf - rowi + 1
        ed = []
        for i in np.nonzero(mask[0])[0]:
            i = int(i)
            if i < rowi or i > rowf:
                continue
            rel = i - rowi
            ed.append((rel if fw else ext - 1 - rel, ord("ACGTN"[min(int(ref[i]), 4)]), ord("ACGTN"[min(int(seq[i]), 4)]), 3))
        if not fw:
            ed = ed[::-1]
        tl, tr = rowi, rdlen - 1 - rowf
        return rc, Aln(tidx, refoff + rowi, fw, int(out[0]["score"]), rdlen, ed, int(out[0]["ns"]), int(out[0]["refns"]), False,
                       tl if fw else tr, tr if fw else tl)

    def dp(self, codes, quals, fw, tidx, rect, minsc, nceil):
        probs = np.zeros(1, dtype=DP_PROBLEM)
        probs[0] = (0, int(fw), tidx, rect.refl, rect.refr, rect.triml, rect.corel, rect.corer, minsc, nceil, 0)
        max_cands = 16384 if self.local else 1024
        summ, cands, alns, ops = self.gpu.dp_extend(self._batch(codes, quals), probs, max_cands=max_cands, max_alns=64,
                                                    max_ops=len(codes) + 80)
        s = summ[0]
        if int(s["flags"]):
            raise RuntimeError(f"bt2g_dp_extend overflow flags {int(s['flags'])}")
        out = dict(found=int(s["found"]), best=int(s["best"]), alns=[], attempts=[])
        if not out["found"]:
            return out
        by_cand = {}
        for k in range(int(s["naln"])):
            a = alns[0][k]
            ed = ops_to_edits(ops[0][k], int(a["nops"]), codes, bool(fw), int(a["row0"]), int(a["trim_end"]))
            t5, t3 = (int(a["trim_beg"]), int(a["trim_end"])) if fw else (int(a["trim_end"]), int(a["trim_beg"]))
            out["alns"].append(dict(score=int(a["score"]), ns=int(a["ns"]), gaps=int(a["gaps"]), refoff=int(rect.refl) + int(a["col0"]),
                                    trim5=t5, trim3=t3, fw=int(fw), edits=[tuple(e) for e in ed]))
            by_cand[int(a["cand_idx"])] = k
        for ci in range(int(s["ncand"])):
            f = int(cands[0][ci]["fate"])
            if f == CAND_SUCCEEDED:
                out["attempts"].append((int(cands[0][ci]["score"]), by_cand[ci]))
            elif f == CAND_FAILED:
                out["attempts"].append((int(cands[0][ci]["score"]), -1))
        return out
