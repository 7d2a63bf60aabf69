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

    @classmethod
    def default(cls, local=False):
        return cls(match_bonus=2 if local else 0, local=local)

    def score_min(self) -> SimpleFunc:
        return SimpleFunc(SIMPLE_FUNC_LOG, 20.0, 8.0) if self.local else SimpleFunc(SIMPLE_FUNC_LINEAR, -0.6, -0.6)

    def n_ceil_func(self) -> SimpleFunc:
        return SimpleFunc(SIMPLE_FUNC_LINEAR, 0.0, 0.15, 0.0, float("inf"))

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
