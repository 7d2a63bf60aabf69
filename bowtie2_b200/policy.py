"""Host-side policy arithmetic around the hot path (pure integer/float restatements).

Each function cites the reference code it mirrors; tests/test_policy.py pins them against the
reference itself (oracle/_ref glue).  Floating point follows the reference's types: SimpleFunc
uses double and truncates toward zero when cast (simple_func.h:89-111).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

SIMPLE_FUNC_CONST, SIMPLE_FUNC_LINEAR, SIMPLE_FUNC_SQRT, SIMPLE_FUNC_LOG = 1, 2, 3, 4


@dataclass
class SimpleFunc:
    """f(x) = max(I, min(X, C + L * g(x))) (simple_func.h:40-125)."""
    type: int
    C: float
    L: float
    I: float = -float("inf")
    X: float = float("inf")

    def f(self, x: float) -> float:
        if self.type == SIMPLE_FUNC_CONST:
            g = 0.0
        elif self.type == SIMPLE_FUNC_LINEAR:
            g = x
        elif self.type == SIMPLE_FUNC_SQRT:
            g = math.sqrt(x)
        else:
            g = math.log(x)
        return max(self.I, min(self.X, self.C + self.L * g))

    def fi(self, x: float) -> int:
        return int(self.f(x))        # C cast: truncation toward zero


def _F32(x: float) -> float:
    """a float literal of the reference widened to double"""
    import struct
    return struct.unpack("f", struct.pack("f", x))[0]


@dataclass
class Preset:
    """presets.cpp:37-92 ("%LOCAL%" variants) -> policy string fields."""
    seed_len: int
    ival: SimpleFunc
    dp_fail_streak: int   # -D
    seed_rounds: int      # -R


def preset(name: str, local: bool = False) -> Preset:
    S = SIMPLE_FUNC_SQRT
    if not local:
        table = {
            "very-fast": Preset(22, SimpleFunc(S, 0.0, 2.50), 5, 1),
            "fast": Preset(22, SimpleFunc(S, 0.0, 2.50), 10, 2),
            "sensitive": Preset(22, SimpleFunc(S, 1.0, 1.15), 15, 2),
            "very-sensitive": Preset(20, SimpleFunc(S, 1.0, 0.50), 20, 3),
        }
    else:
        table = {
            "very-fast": Preset(25, SimpleFunc(S, 1.0, 2.00), 5, 1),
            "fast": Preset(22, SimpleFunc(S, 1.0, 1.75), 10, 2),
            "sensitive": Preset(20, SimpleFunc(S, 1.0, 0.75), 15, 2),
            "very-sensitive": Preset(20, SimpleFunc(S, 1.0, 0.50), 20, 3),
        }
    return table[name]


@dataclass
class Scoring:
    """scoring.h:28-84 defaults (bt2_search.cpp:5040)."""
    match_bonus: int = 0
    rdgap_const: int = 5
    rdgap_linear: int = 3
    rfgap_const: int = 5
    rfgap_linear: int = 3
    gapbar: int = 4
    local: bool = False
    mmp_max: int = 6              # --mp MX,MN
    mmp_min: int = 2
    n_pen: int = 1                # --np
    score_min_func: "SimpleFunc" = None   # --score-min
    n_ceil_over: "SimpleFunc" = None      # --n-ceil

    @classmethod
    def default(cls, local=False):
        return cls(match_bonus=2 if local else 0, local=local)

    def mm_penalty(self, q: int) -> int:
        return mm_penalty(q, self.mmp_max, self.mmp_min)

    def score_min(self) -> SimpleFunc:
        if self.score_min_func is not None:
            return self.score_min_func
        # the defaults are FLOAT literals widened to double (scoring.h:50-55: -0.6f = -0.60000002384...), which moves the
        # truncation point for read lengths where 0.6 * (len + 1) is an integer: 109 bp -> -66, not -65
        return SimpleFunc(SIMPLE_FUNC_LOG, 20.0, 8.0) if self.local else SimpleFunc(SIMPLE_FUNC_LINEAR, _F32(-0.6), _F32(-0.6))

    def n_ceil_func(self) -> SimpleFunc:
        if self.n_ceil_over is not None:
            return self.n_ceil_over
        return SimpleFunc(SIMPLE_FUNC_LINEAR, 0.0, _F32(0.15), 0.0, float("inf"))      # scoring.h:61-63: 0.0f, 0.15f

    def read_gap_open(self): return self.rdgap_const + self.rdgap_linear
    def read_gap_extend(self): return self.rdgap_linear
    def ref_gap_open(self): return self.rfgap_const + self.rfgap_linear
    def ref_gap_extend(self): return self.rfgap_linear

    def perfect_score(self, rdlen: int) -> int:
        return rdlen * self.match_bonus

    def min_score(self, rdlen: int) -> int:
        """bt2_search.cpp:3352-3372"""
        m = self.score_min().fi(rdlen)
        if self.local:
            return max(m, 0)
        return min(m, 0)

    def n_ceil(self, rdlen: int) -> int:
        """bt2_search.cpp:3427-3428 (min with read length); SwAligner::initRead uses the raw value."""
        return min(self.n_ceil_func().fi(rdlen), rdlen)

    def n_ceil_raw(self, rdlen: int) -> int:
        return self.n_ceil_func().fi(rdlen)

    def max_read_gaps(self, minsc: int, rdlen: int) -> int:
        """Scoring::maxReadGaps (scoring.cpp:42-66)"""
        sc = rdlen * self.match_bonus
        first, num = True, 0
        while sc >= minsc:
            sc -= self.read_gap_open() if first else self.read_gap_extend()
            first = False
            num += 1
        return num - 1

    def max_ref_gaps(self, minsc: int, rdlen: int) -> int:
        """Scoring::maxRefGaps (scoring.cpp:73-98)"""
        sc = rdlen * self.match_bonus
        first, num = True, 0
        while sc >= minsc:
            sc -= self.match_bonus
            sc -= self.ref_gap_open() if first else self.ref_gap_extend()
            first = False
            num += 1
        return num - 1


def mm_penalty(q: int, mmp_max: int = 6, mmp_min: int = 2) -> int:
    """Scoring::mm for the default quality-aware model (scoring.h: COST_MODEL_QUAL; float arithmetic)"""
    import numpy as _np
    ii = min(max(q, 0), 40)
    frac = _np.float32(ii) / _np.float32(40.0)
    return mmp_min + int(frac * _np.float32(mmp_max - mmp_min))


def seed_interval(ival: SimpleFunc, rdlen: int, both_mates: bool = False) -> int:
    """bt2_search.cpp:3443-3450"""
    v = ival.fi(float(rdlen))
    if both_mates:
        v = int(v * 1.2 + 0.5)
    return max(v, 1)


def n_seeds(rdlen: int, seed_len: int, interval: int, offset: int = 0) -> int:
    """SeedAligner::instantiateSeeds (aligner_seed.cpp:523-526)"""
    n = 1
    if rdlen - offset > seed_len:
        n += (rdlen - offset - seed_len) // interval
    return n


@dataclass
class DPRect:
    refl: int
    refr: int
    refl_pretrim: int
    refr_pretrim: int
    triml: int
    trimr: int
    corel: int
    corer: int
    maxgap: int

    def entirely_trimmed(self) -> bool:
        return self.refr < self.refl


def frame_seed_extension_rect(off: int, rdlen: int, reflen: int, maxrdgap: int, maxrfgap: int, maxns: int,
                              maxhalf: int = 15, trim_to_ref: bool = True):
    """DynProgFramer::frameSeedExtensionRect (dp_framer.cpp:81-129).  Returns (found, DPRect)."""
    # the reference takes these as size_t: a negative count (minsc above the perfect score) wraps to huge
    m64 = (1 << 64) - 1
    maxgap = min(max(maxrdgap & m64, maxrfgap & m64), maxhalf)
    refl = off - 2 * maxgap
    refr = off + (rdlen - 1) + 2 * maxgap
    triml = trimr = 0
    if trim_to_ref:
        maxns = 0
    elif maxns == rdlen:
        maxns -= 1
    if refr >= reflen + maxns:
        trimr = refr - (reflen + maxns - 1)
    if refl < -maxns:
        triml = (-refl) - maxns
    r = DPRect(refl + triml, refr - trimr, refl, refr, triml, trimr, maxgap, maxgap + 2 * maxgap, maxgap)
    return (not r.entirely_trimmed()), r


# ---- paired-end policy (pe.h / pe.cpp) and mate-finding rectangles (dp_framer.cpp:177-361) ------
PE_POLICY_FF, PE_POLICY_RR, PE_POLICY_FR, PE_POLICY_RF = 1, 2, 3, 4
PE_ALS_NORMAL, PE_ALS_OVERLAP, PE_ALS_CONTAIN, PE_ALS_DOVETAIL, PE_ALS_DISCORD = 1, 2, 3, 4, 5


@dataclass
class PairedEndPolicy:
    """PairedEndPolicy (pe.h:169-330) with the program defaults of bt2_search.cpp:350-358
    (-I 0 -X 500 --fr, no flipping, no dovetail, containment and overlap allowed, expand to fit)."""
    pol: int = PE_POLICY_FR
    maxfrag: int = 500
    minfrag: int = 0
    local: bool = False
    flipping_ok: bool = False
    dovetail_ok: bool = False
    contain_ok: bool = True
    olap_ok: bool = True
    expand_to_fit: bool = True

    def flags(self) -> int:
        return (int(self.flipping_ok) | int(self.dovetail_ok) << 1 | int(self.contain_ok) << 2 | int(self.olap_ok) << 3
                | int(self.expand_to_fit) << 4 | int(self.local) << 5)

    def mate_dir(self, is1: bool, fw: bool):
        """pePolicyMateDir (pe.h:130-164) -> (left, mfw)."""
        if self.pol == PE_POLICY_FF:
            return is1 != fw, fw
        if self.pol == PE_POLICY_RR:
            return is1 == fw, fw
        if self.pol == PE_POLICY_FR:
            return (not fw), (not fw)
        return fw, (not fw)

    def other_mate(self, is1: bool, fw: bool, off: int, maxalcols: int, reflen: int, len1: int, len2: int):
        """PairedEndPolicy::otherMate (pe.cpp:161-355) -> None or (oleft, oll, olr, orl, orr, ofw)."""
        oleft, ofw = self.mate_dir(is1, fw)
        alen = len1 if is1 else len2
        maxfrag, minfrag = self.maxfrag, max(self.minfrag, 1)
        if self.expand_to_fit:
            maxfrag = max(maxfrag, len1, len2)
        elif len1 > maxfrag or len2 > maxfrag:
            return None
        if oleft:
            oll = off + alen - maxfrag
            olr = off + alen - minfrag
            orl = oll
            orr = off + maxfrag - 1
            if not self.olap_ok:
                orr = min(orr, off - 1)
                if orr < olr:
                    olr = orr
            elif not self.dovetail_ok:
                orr = min(orr, off + alen - 1)
            elif not self.flipping_ok and maxalcols != -1:
                orr = min(orr, off + alen - 1 + (maxalcols - 1))
        else:
            orr = off + (maxfrag - 1)
            orl = off + (minfrag - 1)
            oll = off + alen - maxfrag
            olr = orr
            if not self.olap_ok:
                oll = max(oll, off + alen)
                if oll > orl:
                    orl = oll
            elif not self.dovetail_ok:
                oll = max(oll, off)
            elif not self.flipping_ok and maxalcols != -1:
                oll = max(oll, off - maxalcols + 1)
        return oleft, oll, olr, orl, orr, ofw

    def classify_pair(self, off1: int, len1: int, fw1: bool, off2: int, len2: int, fw2: bool) -> int:
        """PairedEndPolicy::peClassifyPair (pe.cpp:37-137)."""
        maxfrag = self.maxfrag
        if self.expand_to_fit:
            maxfrag = max(maxfrag, len1, len2)
        minfrag = max(self.minfrag, 1)
        if self.pol in (PE_POLICY_FF, PE_POLICY_RR):
            if fw1 != fw2:
                return PE_ALS_DISCORD
            one_left = fw1 if self.pol == PE_POLICY_FF else not fw1
        else:
            if fw1 == fw2:
                return PE_ALS_DISCORD
            one_left = fw1 if self.pol == PE_POLICY_FR else not fw1
        frag = max(off1 + len1, off2 + len2) - min(off1, off2)
        if frag > maxfrag or frag < minfrag:
            return PE_ALS_DISCORD
        lo1, hi1, lo2, hi2 = off1, off1 + len1 - 1, off2, off2 + len2 - 1
        containment = (lo1 >= lo2 and hi1 <= hi2) or (lo2 >= lo1 and hi2 <= hi1)
        typ = PE_ALS_NORMAL
        olap = False
        if (lo1 <= lo2 and hi1 >= lo2) or (lo1 <= hi2 and hi1 >= hi2) or containment:
            olap = True
            if not self.olap_ok:
                return PE_ALS_DISCORD
            typ = PE_ALS_OVERLAP
        if not olap:
            if (one_left and lo2 < lo1) or (not one_left and lo1 < lo2):
                return PE_ALS_DISCORD
        if containment:
            if not self.contain_ok:
                return PE_ALS_DISCORD
            typ = PE_ALS_CONTAIN
        if (one_left and (hi1 > hi2 or lo2 < lo1)) or (not one_left and (hi2 > hi1 or lo1 < lo2)):
            if not self.dovetail_ok:
                return PE_ALS_DISCORD
            typ = PE_ALS_DOVETAIL
        return typ


def frame_find_mate_rect(anchor_left: bool, ll: int, lr: int, rl: int, rr: int, rdlen: int, reflen: int,
                         maxrdgap: int, maxrfgap: int, maxns: int, maxhalf: int = 15, trim_to_ref: bool = True):
    """DynProgFramer::frameFindMateRect (dp_framer.h:155-197 -> dp_framer.cpp:177-361).  Returns (found, DPRect).
    NB the mate rectangles take maxgap = max(gaps, maxhalf) (dp_framer.cpp:197-198, :310-311), not min."""
    m64 = (1 << 64) - 1
    maxgap = max(maxrdgap & m64, maxrfgap & m64, maxhalf)
    if anchor_left:
        refl = (rl - (rdlen - 1)) - maxgap
        refr = rr + maxgap
    else:
        refl = ll - maxgap
        refr = (lr + (rdlen - 1)) + maxgap
    triml = trimr = 0
    if trim_to_ref:
        maxns = 0
    elif maxns == rdlen:
        maxns -= 1
    if refr >= reflen + maxns:
        trimr = refr - (reflen + maxns - 1)
    if refl < -maxns:
        triml = (-refl) - maxns
    width = refr - refl + 1
    r = DPRect(refl + triml, refr - trimr, refl, refr, triml, trimr, maxgap, width - maxgap - 1, maxgap)
    return (not r.entirely_trimmed()), r


# ---- MAPQ (unique.h:170-392, BowtieMapq2 = the default model) ---------------------------------------
def mapq_v2(best: int, secbest, sc_min: int, sc_perfect: int, monotone: bool) -> int:
    """BowtieMapq2::mapq for a primary alignment whose search was exhaustive or capped (not the 255 case).
    `secbest` is None when there is no second-best score.  For pairs pass the concordant sums and the summed
    minimum / perfect scores.  Thresholds are float literals widened to double, as in the reference."""
    import struct

    def f(x):                      # (double)0.8f
        return struct.unpack("f", struct.pack("f", x))[0]

    diff = max(1, sc_perfect - sc_min)
    best_over = best - sc_min
    if secbest is None:
        table = ((0.8, 42), (0.7, 40), (0.6, 24), (0.5, 23), (0.4, 8), (0.3, 3)) if monotone else \
                ((0.8, 44), (0.7, 42), (0.6, 41), (0.5, 36), (0.4, 28), (0.3, 24))
        for th, v in table:
            if best_over >= diff * f(th):
                return v
        return 0 if monotone else 22
    bestdiff = abs(abs(best) - abs(secbest))
    if monotone:
        for th, hi, lo in ((0.9, 39, 33), (0.8, 38, 27), (0.7, 37, 26), (0.6, 36, 22)):
            if bestdiff >= diff * f(th):
                return hi if best_over == diff else lo
        for th, top, a, va, b, vb, rest in ((0.5, 35, 0.84, 25, 0.68, 16, 5), (0.4, 34, 0.84, 21, 0.68, 14, 4),
                                             (0.3, 32, 0.88, 18, 0.67, 15, 3), (0.2, 31, 0.88, 17, 0.67, 11, 0),
                                             (0.1, 30, 0.88, 12, 0.67, 7, 0)):
            if bestdiff >= diff * f(th):
                if best_over == diff:
                    return top
                if best_over >= diff * f(a):
                    return va
                if best_over >= diff * f(b):
                    return vb
                return rest
        if bestdiff > 0:
            return 6 if best_over >= diff * f(0.67) else 2
        return 1 if best_over >= diff * f(0.67) else 0
    for th, v in ((0.9, 40), (0.8, 39), (0.7, 38), (0.6, 37)):
        if bestdiff >= diff * f(th):
            return v
    for th, top, va, rest in ((0.5, 35, 25, 20), (0.4, 34, 21, 19), (0.3, 33, 18, 16), (0.2, 32, 17, 12), (0.1, 31, 14, 9)):
        if bestdiff >= diff * f(th):
            if best_over == diff:
                return top
            return va if best_over >= diff * f(0.5) else rest
    if bestdiff > 0:
        return 11 if best_over >= diff * f(0.5) else 2
    return 1 if best_over >= diff * f(0.5) else 0


# ---- per-read pseudo-randomness (random_source.h:32-180, pat.cpp:45-82) ----------------------------
class RandomSource:
    """The reference's linear congruential generator (a = 1664525, c = 1013904223), including the bit-slicing
    helpers nextU2 / nextBool that re-use bits of the last state."""
    A, Cc = 1664525, 1013904223

    def __init__(self, seed: int = 0):
        self.init(seed)

    def init(self, seed: int = 0):
        self.last = seed & 0xffffffff
        self.last_off = 30

    def next_u32(self) -> int:
        self.last = (self.A * self.last + self.Cc) & 0xffffffff
        ret = self.last >> 16
        self.last = (self.A * self.last + self.Cc) & 0xffffffff
        ret ^= self.last
        self.last_off = 0
        return ret

    def next_u2(self) -> int:
        if self.last_off > 30:
            self.next_u32()
        ret = (self.last >> self.last_off) & 3
        self.last_off += 2
        return ret

    def next_bool(self) -> int:
        if self.last_off > 31:
            self.next_u32()
        ret = (self.last >> self.last_off) & 1
        self.last_off += 1
        return ret

    def next_float_bits(self) -> int:
        import struct
        import numpy as _np
        f = _np.float32(self.next_u32()) / _np.float32(0xffffffff)
        return struct.unpack("I", struct.pack("f", float(f)))[0]


def gen_rand_seed(codes, quals, name: str, seed: int = 0) -> int:
    """genRandSeed (pat.cpp:45-82): the per-read RNG seed from the read's bases, qualities (ASCII) and name."""
    rseed = ((seed + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xffffffff
    for i, p in enumerate(codes):
        rseed ^= (int(p) << ((i & 15) << 1)) & 0xffffffff
    for i, p in enumerate(quals):
        rseed ^= (int(p) << ((i & 3) << 3)) & 0xffffffff
    for i, ch in enumerate(name.encode()):
        if ch == ord("/"):
            break
        rseed ^= (ch << ((i & 3) << 3)) & 0xffffffff
    return rseed & 0xffffffff


def rank_seed_hits(nelt_fw, nelt_rc, rnd: RandomSource, all_hits: bool = False):
    """SeedResults::rankSeedHits (aligner_seed.h:1019-1080): order in which (offset index, fw) seed hits are
    extended -- ascending by number of BW elements, the scan start and the strand order drawn from the read's RNG.
    nelt_*[i] = elements of the hit at offset index i (0 = none)."""
    num = len(nelt_fw)
    out = []
    if all_hits:
        for i in range(1, num):
            for fw in (True, False):
                if (nelt_fw if fw else nelt_rc)[i] > 0:
                    out.append((i, fw))
        if num and nelt_fw[0] > 0:
            out.append((0, True))
        if num and nelt_rc[0] > 0:
            out.append((0, False))
        return out
    nonz = sum(1 for x in nelt_fw if x > 0) + sum(1 for x in nelt_rc if x > 0)
    sorted_fw, sorted_rc = [False] * num, [False] * num
    while len(out) < nonz:
        minsz, minidx, minfw = 0xffffffff, 0, True
        rb = rnd.next_bool()
        for fwi in (0, 1):
            fw = fwi == (1 if rb else 0)
            rrs, srt = (nelt_fw, sorted_fw) if fw else (nelt_rc, sorted_rc)
            i = rnd.next_u32() % num
            for _ in range(num):
                if rrs[i] > 0 and not srt[i] and rrs[i] < minsz:
                    minsz, minidx, minfw = rrs[i], i, fw
                i += 1
                if i == num:
                    i = 0
        (sorted_fw if minfw else sorted_rc)[minidx] = True
        out.append((minidx, minfw))
    return out
