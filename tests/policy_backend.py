"""The CPU oracle as a backend of bowtie2_b200.policy_engine (test infrastructure: the product backend issues the
same calls through include/bt2g.h)."""
import numpy as np

from bowtie2_b200 import policy
from bowtie2_b200.policy_engine import Aln
from oracle_lib import extend_both, oracle_dp, oracle_one_mm, oracle_ungapped


class OracleBackend:
    def __init__(self, O, off_size=4, local=False, scoring=None):
        self.O = O
        self.off_size = off_size
        self.local = local
        self.scoring = scoring                     # policy.Scoring with non-default penalties, or None

    def _sc(self, fn, *args, **kw):
        """run one oracle call under this backend's scoring scheme (the override is module state: always restored)"""
        import oracle_lib
        oracle_lib.SCORING_OVERRIDE = self.scoring
        try:
            return fn(*args, **kw)
        finally:
            oracle_lib.SCORING_OVERRIDE = None

    def exact_sweep(self, codes, nofw=False, norc=False):
        return self.O.exact_sweep(codes, nofw, norc)

    def one_mm(self, codes, quals, minsc, nofw, norc):
        return self._sc(oracle_one_mm, self.O, self.local, codes, quals, minsc, nofw, norc)

    def seed_search(self, codes, quals, seed_len, interval, offset, nofw=False, norc=False):
        n = max(1, policy.n_seeds(len(codes), seed_len, interval, offset))
        cnt, out = self.O.seed_search(codes, seed_len, interval, offset, n + 2, nofw, norc, quals=quals)
        return out[:, :cnt, :]

    def extend(self, codes, fw, rdoff, seedlen, rng4):
        return extend_both(self.O, codes, fw, rdoff, seedlen, rng4)

    def resolve(self, row):
        return int(self.O.get_offset(np.array([row], dtype=np.uint64))[0])

    def joined_to_text(self, qlen, off, reject):
        return self.O.joined_to_text(int(qlen), int(off), int(reject))

    def count_ref_ns(self, tidx, off, extent):
        return int((self.O.get_stretch(tidx, off, extent) > 3).sum())

    def ungapped(self, codes, quals, fw, tidx, refoff, tlen, minsc):
        rc, d = self._sc(oracle_ungapped, self.O, self.local, codes, quals, fw, tidx, refoff, tlen, 0, minsc)
        if rc != 1:
            return rc, None
        rdlen = len(codes)
        ref = self.O.get_stretch(tidx, refoff, rdlen)
        seq = codes if fw else np.array([4 if c > 3 else 3 - c for c in codes[::-1]], dtype=np.uint8)
        rowi, rowf = d["rowi"], d["rowf"]                     # aligned rows in reference orientation (local mode trims)
        ext = rowf - rowi + 1
        ed = []
        for i in np.nonzero(d["mask"])[0]:
            i = int(i)
            if i < rowi or i > rowf:
                continue
            rel = i - rowi
            pos = rel if fw else ext - 1 - rel
            ed.append((pos, ord("ACGTN"[min(int(ref[i]), 4)]), ord("ACGTN"[min(int(seq[i]), 4)]), 3))
        if not fw:
            ed = ed[::-1]
        tl, tr = rowi, rdlen - 1 - rowf
        return rc, Aln(tidx, refoff + rowi, fw, d["score"], rdlen, ed, d["ns"], d["refns"], False, tl if fw else tr, tr if fw else tl)

    def dp(self, codes, quals, fw, tidx, rect, minsc, nceil):
        return self._sc(oracle_dp, self.O, self.local, codes, quals, fw, tidx, rect, minsc, nceil, max_cands=65536, max_alns=64,
                        max_edits=16384, attempts=True)
