"""The GPU primitives of include/bt2g.h as the backend of the exact search policy (policy_engine / policy_waves).

`GpuBatchBackend.batch(name, requests, ids)` answers all requests of one primitive with ONE call of the matching entry
point (the entry points take arrays; the DP is chunked to bound its output buffers); `GpuBackend` is the per-item view
of the same code.  Conventions are those the -m gpu parity tests establish for each entry point (test_fm_gpu,
test_onemm, test_extend, test_ungapped, test_dp_gpu), where the results are pinned bit-exact against the CPU oracle.
The DP's backtrace attempts come from the per-candidate fates bt2g_dp_extend returns (BT2G_CAND_FAILED / _SUCCEEDED).

tests/test_exact_gpu.py: engine over this backend == golden SAM, per item and in waves."""
import numpy as np

from . import policy
from .lib import DP_PROBLEM, UNGAPPED_PROBLEM, Bt2Gpu, ReadBatch, ops_to_edits
from .policy_engine import Aln

CAND_FAILED, CAND_SUCCEEDED = 2, 3
DNA = "ACGTN"


def _u8(x):
    return np.asarray(x, dtype=np.uint8)


class GpuBatchBackend:
    DP_CHUNK = 4096

    def __init__(self, gpu: Bt2Gpu, local: bool = False, sc: "policy.Scoring" = None):
        self.gpu = gpu
        self.local = local
        if hasattr(gpu, "set_scoring_policy"):
            gpu.set_scoring_policy(sc or policy.Scoring.default(local), local)
        else:                                        # the stand-in device of the CPU test-suite
            gpu.set_scoring(local=local)
        self.off_size = int(gpu.info()["off_size"])
        self._rows = {}                              # id -> BW row of the read's last resolve request

    def batch(self, name, requests, ids):
        return getattr(self, "_b_" + name)(requests, ids)

    # ---- helpers
    @staticmethod
    def _grouped(requests, keyf):
        groups = {}
        for k, a in enumerate(requests):
            groups.setdefault(keyf(a), []).append(k)
        return groups

    # ---- FM primitives
    def _b_exact_sweep(self, requests, ids):
        out = [None] * len(requests)
        for (nofw, norc), idx in self._grouped(requests, lambda a: (bool(a[1]), bool(a[2]))).items():
            mine, ee = self.gpu.exact_sweep(ReadBatch.from_list([_u8(requests[k][0]) for k in idx]), nofw, norc)
            for j, k in enumerate(idx):
                tb = [int(x) for x in ee[j]]
                out[k] = (max(0, tb[1] - tb[0]) + max(0, tb[3] - tb[2]), [int(mine[j][0]), int(mine[j][1])], tb)
        return out

    def _b_one_mm(self, requests, ids):
        batch = ReadBatch.from_list([_u8(a[0]) for a in requests], [_u8(a[1]) for a in requests])
        minsc = np.array([int(a[2]) for a in requests], dtype=np.int32)
        mask = np.array([(0 if a[3] else 1) | (0 if a[4] else 2) for a in requests], dtype=np.uint8)
        hits, counts = self.gpu.one_mm(batch, minsc, mask, max_hits=64)
        out = []
        for k in range(len(requests)):
            lst = []
            for task in range(4):
                for h in hits[k, task, :counts[k, task]]:
                    lst.append((int(h["top"]), int(h["bot"]), int(h["pos"]), ord(DNA[int(h["chr"])]), ord(DNA[int(h["qchr"])]),
                                int(h["score"]), int(task < 2)))
            out.append(lst)
        return out

    def _b_seed_search(self, requests, ids):
        # args: codes, quals, seed_len, interval, offset, nofw, norc
        out = [None] * len(requests)
        for (L, nofw, norc), idx in self._grouped(requests, lambda a: (int(a[2]), bool(a[5]), bool(a[6]))).items():
            rq = [requests[k] for k in idx]
            ns_max = max(max(1, policy.n_seeds(len(a[0]), L, int(a[3]), int(a[4]))) for a in rq) + 2
            batch = ReadBatch.from_list([_u8(a[0]) for a in rq], [_u8(a[1]) for a in rq])
            res, ns = self.gpu.seed_search(batch, L, np.array([int(a[3]) for a in rq], dtype=np.int32),
                                           np.array([int(a[4]) for a in rq], dtype=np.int32), ns_max, nofw, norc)
            for j, k in enumerate(idx):
                out[k] = res[j][:, :int(ns[j]), :].copy()
        return out

    def _b_extend(self, requests, ids):
        # args: codes, fw, rdoff, seedlen, (topf, botf, topb, botb): one seed per request at offset rdoff
        out = [None] * len(requests)
        for L, idx in self._grouped(requests, lambda a: int(a[3])).items():
            rq = [requests[k] for k in idx]
            ranges = np.zeros((len(rq), 2, 1, 4), dtype=np.uint64)
            for j, a in enumerate(rq):
                ranges[j, 0 if a[1] else 1, 0] = [int(x) for x in a[4]]
            batch = ReadBatch.from_list([_u8(a[0]) for a in rq])
            ext = self.gpu.extend_exact(batch, L, np.array([max(1, len(a[0])) for a in rq], dtype=np.int32),
                                        np.array([int(a[2]) for a in rq], dtype=np.int32), 1, ranges)
            for j, k in enumerate(idx):
                e = ext[j, 0 if rq[j][1] else 1, 0]
                out[k] = (int(e[0]), int(e[1]))
        return out

    def _b_resolve(self, requests, ids):
        rows = np.array([int(a[0]) for a in requests], dtype=np.uint64)
        for i, r in zip(ids, rows):
            self._rows[i] = int(r)
        joined, *_ = self.gpu.resolve(rows, 1, False)
        return [int(x) for x in joined]

    def _b_joined_to_text(self, requests, ids):
        # args: qlen, joined offset, reject straddlers; the entry point resolves and converts in one pass from the row
        out = [None] * len(requests)
        for reject, idx in self._grouped(requests, lambda a: bool(a[2])).items():
            rows = np.array([self._rows[ids[k]] for k in idx], dtype=np.uint64)
            qlen = np.array([int(requests[k][0]) for k in idx], dtype=np.uint32)
            joined, tidx, textoff, tlen, flags = self.gpu.resolve(rows, qlen, reject)
            for j, k in enumerate(idx):
                assert int(joined[j]) == int(requests[k][1])
                f = int(flags[j])
                out[k] = (not ((f >> 1) & 1), int(tidx[j]), int(textoff[j]), int(tlen[j]), f & 1)
        return out

    def _b_count_ref_ns(self, requests, ids):
        stride = max(int(a[2]) for a in requests)
        s = self.gpu.get_stretch([a[0] for a in requests], [a[1] for a in requests], [a[2] for a in requests], stride)
        return [int((s[k][:int(a[2])] > 3).sum()) for k, a in enumerate(requests)]

    # ---- ungapped and gapped extension
    def _b_ungapped(self, requests, ids):
        # args: codes, quals, fw, tidx, refoff, tlen, minsc
        batch = ReadBatch.from_list([_u8(a[0]) for a in requests], [_u8(a[1]) for a in requests])
        probs = np.zeros(len(requests), dtype=UNGAPPED_PROBLEM)
        for k, a in enumerate(requests):
            probs[k] = (k, int(a[2]), a[3], a[4], a[5], a[6], 0)
        res, mask = self.gpu.ungapped(batch, probs)
        hit = [k for k in range(len(requests)) if int(res[k]["status"]) == 1]
        refs = {}
        if hit:
            stride = max(len(requests[k][0]) for k in hit)
            st = self.gpu.get_stretch([requests[k][3] for k in hit], [requests[k][4] for k in hit], [len(requests[k][0]) for k in hit], stride)
            refs = {k: st[j] for j, k in enumerate(hit)}
        out = []
        for k, a in enumerate(requests):
            rc = int(res[k]["status"])
            if rc != 1:
                out.append((rc, None))
                continue
            codes, fw, tidx, refoff = _u8(a[0]), bool(a[2]), a[3], a[4]
            rdlen = len(codes)
            seq = codes if fw else np.array([4 if c > 3 else 3 - c for c in codes[::-1]], dtype=np.uint8)
            rowi, rowf = int(res[k]["rowi"]), int(res[k]["rowf"])
            ext = rowf - rowi + 1
            ed = []
            for i in np.nonzero(mask[k][:rdlen])[0]:
                i = int(i)
                if i < rowi or i > rowf:
                    continue
                rel = i - rowi
                ed.append((rel if fw else ext - 1 - rel, ord(DNA[min(int(refs[k][i]), 4)]), ord(DNA[min(int(seq[i]), 4)]), 3))
            if not fw:
                ed = ed[::-1]
            tl, tr = rowi, rdlen - 1 - rowf
            out.append((rc, Aln(tidx, refoff + rowi, fw, int(res[k]["score"]), rdlen, ed, int(res[k]["ns"]), int(res[k]["refns"]), False,
                                tl if fw else tr, tr if fw else tl)))
        return out

    def _b_dp(self, requests, ids):
        # args: codes, quals, fw, tidx, rect, minsc, nceil
        out = []
        for c0 in range(0, len(requests), self.DP_CHUNK):
            rq = requests[c0:c0 + self.DP_CHUNK]
            batch = ReadBatch.from_list([_u8(a[0]) for a in rq], [_u8(a[1]) for a in rq])
            probs = np.zeros(len(rq), dtype=DP_PROBLEM)
            for k, a in enumerate(rq):
                r = a[4]
                probs[k] = (k, int(a[2]), a[3], r.refl, r.refr, r.triml, r.corel, r.corer, a[5], a[6], 0)
            max_cands = 16384 if self.local else 1024
            max_alns = 16
            summ, cands, alns, ops = self.gpu.dp_extend(batch, probs, max_cands=max_cands, max_alns=max_alns,
                                                        max_ops=int(batch.lengths().max()) + 80)
            for k, a in enumerate(rq):
                s = summ[k]
                if int(s["flags"]):
                    # rare: more alignments / candidates than the batch buffers hold -> this problem alone, larger buffers
                    p1 = probs[k:k + 1].copy()
                    p1["read_idx"] = 0
                    s1, c1, a1, o1 = self.gpu.dp_extend(ReadBatch.from_list([_u8(a[0])], [_u8(a[1])]), p1, max_cands=65536, max_alns=128,
                                                        max_ops=len(a[0]) + 80)
                    if int(s1[0]["flags"]):
                        raise RuntimeError(f"bt2g_dp_extend overflow flags {int(s1[0]['flags'])}")
                    out.append(self._dp_result(s1[0], c1[0], a1[0], o1[0], a))
                else:
                    out.append(self._dp_result(s, cands[k], alns[k], ops[k], a))
        return out

    @staticmethod
    def _dp_result(s, cands, alns, ops, a):
        codes, fw, rect = _u8(a[0]), bool(a[2]), a[4]
        out = dict(found=int(s["found"]), best=int(s["best"]), alns=[], attempts=[])
        if not out["found"]:
            return out
        by_cand = {}
        for k in range(int(s["naln"])):
            al = alns[k]
            ed = ops_to_edits(ops[k], int(al["nops"]), codes, fw, int(al["row0"]), int(al["trim_end"]))
            t5, t3 = (int(al["trim_beg"]), int(al["trim_end"])) if fw else (int(al["trim_end"]), int(al["trim_beg"]))
            out["alns"].append(dict(score=int(al["score"]), ns=int(al["ns"]), gaps=int(al["gaps"]), refoff=int(rect.refl) + int(al["col0"]),
                                    trim5=t5, trim3=t3, fw=int(fw), edits=[tuple(e) for e in ed]))
            by_cand[int(al["cand_idx"])] = k
        for ci in range(int(s["ncand"])):
            f = int(cands[ci]["fate"])
            if f == CAND_SUCCEEDED:
                out["attempts"].append((int(cands[ci]["score"]), by_cand[ci]))
            elif f == CAND_FAILED:
                out["attempts"].append((int(cands[ci]["score"]), -1))
        return out


class GpuBackend:
    """per-item view (one entry-point call per request): policy_engine.PolicyEngine(GpuBackend(gpu), ...)"""

    def __init__(self, gpu: Bt2Gpu, local: bool = False, sc: "policy.Scoring" = None):
        self.bb = GpuBatchBackend(gpu, local, sc)
        self.off_size = self.bb.off_size
        self.local = local

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return lambda *args: self.bb.batch(name, [args], [0])[0]
