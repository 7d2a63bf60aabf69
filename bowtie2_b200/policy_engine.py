"""Sequential search policy of the reference for ONE unpaired end-to-end read, replayed exactly.

The batched pipeline (pipeline.cu) speculates: it extends every plausible seed hit at once and keeps the best
result.  The reference instead walks a per-read, RNG-driven sequence (multiseedSearchWorker, bt2_search.cpp:3101-4250;
SwDriver::extendSeeds, aligner_sw_driver.cpp:921-1500) whose decisions -- which BW rows get resolved, when the
-M ceiling or a failure streak stops the search, how ties are broken -- decide MAPQ, XS:i and, in repeats, the
reported locus.  This module restates that control flow over an abstract `backend` that supplies the hot-path
primitives (exact sweep, 1-mismatch search, seed search, seed extension, offset resolution, ungapped and gapped
DP), so that records can be made byte-identical to the reference program's.  The test-suite plugs in the CPU oracle
to pin the control flow against the reference's SAM; a product backend issues the same calls through include/bt2g.h.

Scope: unpaired reads, end-to-end mode, default reporting (-M, no -k/-a), no --local, no mates.  Everything
numeric that the reference does in float/double is done in the same type here.

Every RNG draw of the reference on this path is reproduced, in order:
  rnd.init(genRandSeed)                         bt2_search.cpp:3437-3440, pat.cpp:45-82
  eeSaTups: strand order, range trimming        aligner_sw_driver.cpp:101-131, 205-221 (sort1mmEe)
  rankSeedHits                                   aligner_seed.h:1019-1080
  prioritizeSATupsRands: RowSampler, Random1toN  aligner_sw_driver.cpp:690-706
  extendSeeds: Random1toN per range              aligner_sw_driver.cpp:1120
  nextAlignment: reseed per backtrace attempt    aligner_sw.cpp:794-795, 877 (8-bit) / 879, 932 (16-bit)
  finishRead: selectByScore tie shuffles         aln_sink.cpp:1552-1568
"""
from dataclasses import dataclass, field

import numpy as np

from . import policy
from .policy import RandomSource

MIN_I64 = -(1 << 63)
EXHAUSTED, FULFILLED, PERFECT, SOFT_LIMIT, HARD_LIMIT = 1, 2, 3, 4, 5


def _next_u64(rnd):
    hi = rnd.next_u32()
    return (hi << 32) | rnd.next_u32()


def _next_float(rnd):
    """RandomSource::nextFloat (random_source.h:137-140): float32 division, widened to double by the callers"""
    return float(np.float32(rnd.next_u32()) / np.float32(0xffffffff))


def shuffle_portion(lst, begin, num, rnd):
    """EList::shufflePortion (ds.h:804-815) with 64-bit size_t"""
    if num < 2:
        return
    left = num
    for i in range(begin, begin + num - 1):
        r = _next_u64(rnd) % left
        if r > 0:
            lst[i], lst[i + r] = lst[i + r], lst[i]
        left -= 1


def shuffle_equal_streaks(lst, key, rnd):
    """the streak loop shared by selectByScore (aln_sink.cpp:1552-1568) and sort1mmEe (aligner_seed.h:1223-1245)"""
    streak = 0
    for i in range(1, len(lst)):
        if key(lst[i]) == key(lst[i - 1]):
            if streak == 0:
                streak = 1
            streak += 1
        else:
            if streak > 1:
                shuffle_portion(lst, i - streak, streak, rnd)
            streak = 0
    if streak > 1:
        shuffle_portion(lst, len(lst) - streak, streak, rnd)


class Random1toN:
    """random_util.h:32-160: pseudo-randoms from [0, n) without replacement"""
    SWAPLIST_THRESH, CONVERSION_THRESH = 128, 16

    def __init__(self):
        self.n = self.cur = 0
        self.swaplist = self.converted = False
        self.list, self.seen, self.thresh = [], [], 0

    def init(self, n, without_replacement):
        self.n, self.cur = n, 0
        self.converted = False
        self.swaplist = n < self.SWAPLIST_THRESH or without_replacement
        self.list, self.seen = [], []
        self.thresh = max(self.CONVERSION_THRESH, int(np.float32(0.10) * np.float32(n)))

    def inited(self):
        return self.n > 0

    def done(self):
        return self.inited() and self.cur >= self.n

    def next(self, rnd):
        if self.cur == 0 and not self.converted:
            if self.n == 1:
                self.cur = 1
                return 0
            if self.swaplist:
                self.list = list(range(self.n))
        if self.swaplist:
            r = self.cur + (rnd.next_u32() % (self.n - self.cur))
            if r != self.cur:
                self.list[self.cur], self.list[r] = self.list[r], self.list[self.cur]
            self.cur += 1
            return self.list[self.cur - 1]
        seen_sz = len(self.seen)
        while True:
            rn = rnd.next_u32() % self.n
            if rn not in self.seen[:seen_sz]:
                break
        self.seen.append(rn)
        self.cur += 1
        if len(self.seen) >= self.thresh and self.cur < self.n:
            s = set(self.seen)
            self.list = [j for j in range(self.n) if j not in s]
            self.seen = []
            self.cur = 0
            self.n = len(self.list)
            self.converted = True
            self.swaplist = True
        return rn


class RowSampler:
    """aligner_sw_driver.h:179-256: weighted choice of the next non-small range (double arithmetic)"""

    def __init__(self, sats, lensq=True, szsq=True):
        self.masses, self.elim, self.mass = [], [False] * len(sats), 0.0
        for s in sats:
            num = float(s.nlex + s.nrex + 1)
            if lensq:
                num *= num
            den = float(s.size)
            if szsq:
                den *= den
            self.masses.append(num / den)
            self.mass += self.masses[-1]

    def finished(self, i):
        self.elim[i] = True
        self.mass -= self.masses[i]

    def next(self, rnd):
        rd = _next_float(rnd) * self.mass
        sofar, last = 0.0, None
        for i, m in enumerate(self.masses):
            if not self.elim[i]:
                last = i
                sofar += m
                if rd < sofar:
                    return i
        return last


class IntervalSet:
    """EIvalMergeListBinned as used by SwDriver (seenDiags1_): membership of (ref, strand, offset) in a union of intervals"""

    def __init__(self):
        self.iv = {}

    def add(self, tidx, fw, off, length):
        self.iv.setdefault((tidx, fw), []).append((off, off + length))

    def present(self, tidx, fw, off):
        return any(a <= off < b for a, b in self.iv.get((tidx, fw), ()))


@dataclass
class Aln:
    """what the engine needs of an AlnRes"""
    tidx: int
    refoff: int
    fw: bool
    score: int
    rdlen: int
    edits: list            # reference Edit convention: (pos from the 5' end, chr, qchr, type 1 read gap / 2 ref gap / 3 mismatch)
    ns: int = 0
    refns: int = 0
    exact: bool = False     # came from the end-to-end exact / 1-mismatch search
    trim5: int = 0          # soft-trimmed read characters at the 5' / 3' end (local mode)
    trim3: int = 0

    @property
    def ext(self):          # AlnRes::readExtentRows
        return self.rdlen - self.trim5 - self.trim3

    @property
    def trim_left(self):    # AlnRes::trimmedLeft(true): in reference orientation
        return self.trim5 if self.fw else self.trim3

    @property
    def ref_extent(self):
        return self.ext + sum(e[3] == 1 for e in self.edits) - sum(e[3] == 2 for e in self.edits)


def edits_left_to_right(a: Aln):
    """AlnRes::invertEdits for reverse-strand alignments (edit.cpp:50-78): positions from the left end of the aligned
    extent in reference orientation"""
    if a.fw:
        return [tuple(e) for e in a.edits]
    out = []
    for pos, ch, qch, typ in reversed(a.edits):
        out.append((a.ext - pos - (0 if typ == 1 else 1), ch, qch, typ))
    return out


class RedundantAlns:
    """aligner_result.cpp:929-1030: cells (row, ref column) already covered by a reported alignment"""

    def __init__(self):
        self.cells = set()

    def _walk(self, a: Aln):
        ned = edits_left_to_right(a)
        left = a.refoff
        k = 0
        start = a.trim_left          # the reference compares edit positions (relative to the extent) with rows counted
        n = start + a.ext            # from the untrimmed read start: restated as written
        for i in range(start, n):
            diff = 1
            right = left + 1
            while k < len(ned) and ned[k][0] == i:
                if ned[k][3] == 2:
                    diff = 0
                k += 1
            if i < n - 1:
                k2 = k
                while k2 < len(ned) and ned[k2][0] == i + 1:
                    if ned[k2][3] == 1:
                        right += 1
                    k2 += 1
            for j in range(left, right):
                yield (a.tidx, a.fw, j, i)
            left = right + diff - 1

    def overlap(self, a: Aln):
        return any(c in self.cells for c in self._walk(a))

    def add(self, a: Aln):
        self.cells.update(self._walk(a))


class UnpairedSink:
    """AlnSinkWrap + ReportingState for an unpaired read in -M mode (aln_sink.cpp:60-330, 1395-1452)"""

    def __init__(self, khits=1, mhits=50, mmode=True):
        self.khits, self.mhits, self.mmode = khits, mhits, mmode
        self.alns = []
        self.done = False
        self.exit_m = False
        self.exit_k = False
        self.best = self.best2 = MIN_I64

    def report(self, a: Aln):
        self.alns.append(a)
        if not self.done:                                  # ReportingState::areDone (aln_sink.cpp:305-318)
            if not self.mmode and len(self.alns) >= self.khits:
                self.done = self.exit_k = True
            elif self.mmode and len(self.alns) > self.mhits:
                self.done = self.exit_m = True
        if a.score > self.best:
            self.best2, self.best = self.best, a.score
        elif a.score > self.best2:
            self.best2 = a.score
        return self.done

    def done_with_mate(self):
        return self.done


@dataclass
class SatPos:
    topf: int
    topb: int
    size: int
    key_len: int
    fw: bool
    offidx: int
    rdoff: int
    seedlen: int
    nlex: int = 0
    nrex: int = 0
    orig_size: int = 0

    def sort_key(self):
        # SATuple::operator< (aligner_cache.h:399-407) then SeedPos::operator< (aligner_sw_driver.h:118-128)
        return (self.size, self.topf, self.offidx, self.rdoff, self.seedlen, 0 if self.fw else 1)


@dataclass
class EEHit:
    top: int
    bot: int
    fw: bool
    score: int
    edit: tuple = None      # (pos, chr, qchr) of the single mismatch, reference Edit convention

    def mms(self):
        return 0 if self.edit is None else 1

    def ns(self):
        return int(self.edit is not None and (self.edit[1] == ord("N") or self.edit[2] == ord("N")))

    def refns(self):
        return int(self.edit is not None and self.edit[1] == ord("N"))


@dataclass
class ReadResult:
    aligned: bool = False
    aln: Aln = None
    xs: int = None
    mapq: int = 0
    filtered: str = None         # "NS" / "LN" when the read never entered the search
    n_alns: int = 0
    maxed: bool = False
    secondary: list = None       # -k / -a: the further selected alignments, in report order (FLAG 256, MAPQ 255)
    counters: dict = None        # nExIters / nExDps / nExUgs / nRedundants of the reference's per-read metrics (ZI XD XU YR)


@dataclass
class MateCtx:
    """per-mate state of the worker loop and of SwDriver (one of these is the "anchor" during extendSeeds*)"""
    codes: np.ndarray
    quals: np.ndarray
    name: str
    rdlen: int = 0
    minsc: int = 0
    perfect: int = 0
    nceil: int = 0
    filt: bool = True
    filtered: str = None
    mm1: list = field(default_factory=list)            # SeedResults::mm1Hit_
    ee: list = field(default_factory=list)             # exact end-to-end hits
    ex_ranges: dict = field(default_factory=lambda: {True: [], False: []})   # seedExRangeFw_/Rc_
    seen: "IntervalSet" = field(default_factory=lambda: IntervalSet())           # seenDiags1_/2_
    sh: dict = None                                    # seed hits of the current round


class PolicyEngine:
    def __init__(self, backend, preset="sensitive", seed=0, sc=None, local=False, nofw=False, norc=False,
                 dp_fail_streak=None, seed_rounds=None, seed_len=None, k=None, all_hits=False, mhits=None, ival=None):
        """seed = --seed; nofw / norc = --nofw / --norc; dp_fail_streak / seed_rounds / seed_len = -D / -R / -L on top of the
        preset; k = -k <int> (up to k alignments per read, no -M sampling); all_hits = -a"""
        self.b = backend
        self.off_size = backend.off_size if backend is not None else 4
        self.local = local
        self.pre = policy.preset(preset, local)
        if dp_fail_streak is not None:
            self.pre.dp_fail_streak = dp_fail_streak
        if seed_rounds is not None:
            self.pre.seed_rounds = seed_rounds
        if seed_len is not None:
            self.pre.seed_len = seed_len
        self.sc = sc or policy.Scoring.default(local)
        self.seed = seed
        self.gnofw, self.gnorc = nofw, norc
        # bt2_search.cpp:342-343, 459-492 and the preset's -D / -R
        # default: -M mode (khits 1, mhits 50, bt2_search.cpp:342-343); -k / -a switch -M off (:1772-1774)
        self.all = all_hits
        self.mmode = not (all_hits or k is not None)
        self.khits = (1 << 62) if all_hits else (k if k is not None else 1)
        self.mhits = (mhits if mhits is not None else 50) if self.mmode else (1 << 62)     # -M <n>
        if ival is not None:                           # -i <func>
            self.pre.ival = ival
        self.maxhalf = 15
        self.max_iters, self.max_ug, self.max_dp = 400, 300, 300
        self.streak = self.pre.dp_fail_streak
        self.max_mate_streak = 10
        if all_hits:                                   # bt2_search.cpp:3459-3473
            self.max_iters = self.max_ug = self.max_dp = self.streak = self.max_mate_streak = 1 << 62
        elif self.khits > 1:
            self.streak += (self.khits - 1) * 10
            self.max_mate_streak += (self.khits - 1) * 10
            self.max_iters += (self.khits - 1) * 20
            self.max_ug += (self.khits - 1) * 20
            self.max_dp += (self.khits - 1) * 20
        self.n_seed_rounds = self.pre.seed_rounds
        self.tighten = 3
        self.seed_boost_thresh = 300
        self.nsm = 5

    # ------------------------------------------------------------------------------------------------- one read
    def _drive(self, gen):
        """run a step generator to completion against self.b (one backend call per request)"""
        try:
            req = next(gen)
            while True:
                req = gen.send(getattr(self.b, req[0])(*req[1]))
        except StopIteration as e:
            return e.value

    def align_read(self, codes, quals, name) -> ReadResult:
        return self._drive(self.read_steps(codes, quals, name))

    def read_steps(self, codes, quals, name):
        """generator form of the policy: yields (primitive name, argument tuple) requests, receives their results, and
        returns the ReadResult.  The wave scheduler (policy_waves.py) interleaves many of these so that each primitive
        runs as one batched call per wave."""
        sc = self.sc
        rdlen = len(codes)
        res = ReadResult()
        codes = np.asarray(codes, dtype=np.uint8)
        quals = np.asarray(quals, dtype=np.uint8)
        ns_in_read = int((codes > 3).sum())
        if rdlen < 2 or rdlen <= 0:
            res.filtered = "LN"
            return res
        if ns_in_read > sc.n_ceil(rdlen):
            res.filtered = "NS"
            return res
        if sc.perfect_score(rdlen) < sc.min_score(rdlen):        # Scoring::scoreFilter (bt2_search.cpp:3385)
            res.filtered = "SC"
            return res
        self.cur = MateCtx(codes, quals, name, rdlen, sc.min_score(rdlen), sc.perfect_score(rdlen), sc.n_ceil(rdlen))
        rnd = self.rnd = RandomSource(policy.gen_rand_seed(codes, quals, name, self.seed))
        interval = policy.seed_interval(self.pre.ival, rdlen, False)
        # per-read state of SwDriver (nextRead) and of the sink
        self.red = RedundantAlns()
        self.sink = UnpairedSink(self.khits, self.mhits, self.mmode)
        self.n_iters = self.n_dps = self.n_ugs = self.n_red = 0
        done = False
        # ---- exact end-to-end (bt2_search.cpp:3493-3690)
        nofw, norc = self.gnofw, self.gnorc
        nelt, mined, tb = yield ("exact_sweep", (codes, nofw, norc,))
        minedfw, minedrc = int(mined[0]), int(mined[1])
        if nelt > 0:
            ee = []
            if tb[1] > tb[0]:
                ee.append(EEHit(int(tb[0]), int(tb[1]), True, self.cur.perfect))
            if tb[3] > tb[2]:
                ee.append(EEHit(int(tb[2]), int(tb[3]), False, self.cur.perfect))
            ret = yield from self.extend_seeds(None, ee)
            done = self._after_extend(ret, done)
        # ---- 1-mismatch end-to-end (bt2_search.cpp:3692-3875)
        if not done:
            yfw, yrc = minedfw <= 1 and not nofw, minedrc <= 1 and not norc
            if yfw or yrc:
                hits = yield ("one_mm", (codes, quals, self.cur.minsc, not yfw, not yrc,))
                self.cur.mm1 = [EEHit(int(h[0]), int(h[1]), bool(h[6]), int(h[5]), (int(h[2]), int(h[3]), int(h[4]))) for h in hits]
                if self.cur.mm1 and not self.sink.done_with_mate():
                    ret = yield from self.extend_seeds(None, [])
                    self.cur.mm1 = []                           # clear1mmE2eHits (bt2_search.cpp:3839)
                    done = self._after_extend(ret, done)
                elif self.cur.mm1:
                    done = True
        # ---- seed rounds (bt2_search.cpp:3876-4150)
        nrounds = min(self.n_seed_rounds, interval)
        L = self.pre.seed_len
        for roundi in range(self.n_seed_rounds):
            if done or self.sink.done_with_mate():
                done = True
                break
            if roundi >= nrounds or interval <= roundi:
                continue
            offset = (interval * roundi) // nrounds
            if offset > 0 and L + offset > rdlen:
                continue
            hits = yield ("seed_search", (codes, quals, min(L, rdlen), interval, offset, nofw, norc,))
            if hits is None:                                   # no seed could be instantiated
                done = True
                break
            nelt_fw = [max(0, int(h[1]) - int(h[0])) for h in hits[0]]
            nelt_rc = [max(0, int(h[1]) - int(h[0])) for h in hits[1]]
            nonz = sum(x > 0 for x in nelt_fw) + sum(x > 0 for x in nelt_rc)
            if nonz == 0:
                done = True
                break
            ranks = policy.rank_seed_hits(nelt_fw, nelt_rc, rnd, self.all)
            sh = dict(hits=hits, ranks=ranks, interval=interval, offset=offset, seedlen=min(L, rdlen), nonz=nonz,
                      nelt=sum(nelt_fw) + sum(nelt_rc))
            ret = yield from self.extend_seeds(sh, [])
            done = self._after_extend(ret, done, check_perfect=False)
            if not done and sh["nelt"] // nonz < self.seed_boost_thresh:
                done = True
        return self.finish_read(res)

    def _after_extend(self, ret, done, check_perfect=True):
        if ret == FULFILLED:
            if self.sink.done_with_mate():
                done = True
        elif ret in (PERFECT, HARD_LIMIT):
            done = True
        if check_perfect and not done and self.cur.minsc == self.cur.perfect:
            done = True
        return done

    # ------------------------------------------------------------------------------------------- eeSaTups
    def _ee_sa_tups(self, ee_exact, maxelt):
        """aligner_sw_driver.cpp:66-290 -> list of (SatPos, EEHit, Random1toN)"""
        rnd = self.rnd
        out = []
        nelt = 0
        done = False
        tot = sum(h.bot - h.top for h in ee_exact)
        if tot > 0:
            fw_first = True
            fwsz = sum(h.bot - h.top for h in ee_exact if h.fw)
            rn = (rnd.next_u32() if self.off_size == 4 else _next_u64(rnd)) % tot
            if rn >= fwsz:
                fw_first = False
            for fwi in range(2):
                if done:
                    break
                fw = (fwi == 0) == fw_first
                hit = next((h for h in ee_exact if h.fw == fw), None)
                if hit is None:
                    continue
                nelt, done = self._ee_add(out, hit, nelt, maxelt, done)
        if not done and self.cur.mm1:
            # EList::sort = std::sort; the lists here are short enough for its insertion-sort (stable) regime
            self.cur.mm1.sort(key=lambda h: -h.score)
            shuffle_equal_streaks(self.cur.mm1, lambda h: h.score, rnd)
            for hit in self.cur.mm1:
                if done:
                    break
                nelt, done = self._ee_add(out, hit, nelt, maxelt, done)
        return out, nelt

    def _ee_add(self, out, hit, nelt, maxelt, done):
        rnd = self.rnd
        tops, bots = [hit.top, 0], [hit.bot, 0]
        width = hit.bot - hit.top
        if width <= 0:
            return nelt, done
        if nelt + width > maxelt:
            trim = (nelt + width) - maxelt
            rn = (rnd.next_u32() if self.off_size == 4 else _next_u64(rnd)) % width
            newwidth = width - trim
            if hit.top + rn + newwidth > hit.bot:
                tops[0], bots[0] = hit.top + rn, hit.bot
                tops[1], bots[1] = hit.top, hit.top + newwidth - (bots[0] - tops[0])
            else:
                tops[0] = hit.top + rn
                bots[0] = tops[0] + newwidth
        for i in range(2):
            if done or bots[i] <= tops[i]:
                break
            w = bots[i] - tops[i]
            sp = SatPos(tops[i], 0, w, self.cur.rdlen, hit.fw, 0, 0, self.cur.rdlen, orig_size=w)
            r = Random1toN()
            r.init(w, self.all)
            out.append((sp, hit, r))
            nelt += w
            if nelt >= maxelt:
                done = True
        return nelt, done

    # ------------------------------------------------------------------------------- prioritizeSATupsRands
    def _prioritize(self, sh, maxelt):
        """aligner_sw_driver.cpp:490-725 -> (list of (SatPos, None, Random1toN), nelt)"""
        rnd = self.rnd
        sats = []
        nelt = 0
        for offidx, fw in sh["ranks"]:
            h = sh["hits"][0 if fw else 1][offidx]
            topf, botf, topb, botb = (int(x) for x in h)
            sz = botf - topf
            rdoff = sh["offset"] + offidx * sh["interval"]
            seedlen = sh["seedlen"]
            nelt += sz
            rng = self.cur.ex_ranges[fw]
            if any(p5 <= rdoff and p5 + ln >= rdoff + seedlen and sz <= rsz for p5, ln, rsz in rng):
                nelt -= sz
                continue
            sp = SatPos(topf, topb, sz, seedlen, fw, offidx, rdoff, seedlen, orig_size=sz)
            sp.nlex, sp.nrex = yield ("extend", (self.cur.codes, fw, rdoff, seedlen, (topf, botf, topb, botb),))
            if sp.nlex > 0 or sp.nrex > 0:
                rng.append((rdoff - (sp.nlex if fw else sp.nrex), seedlen + sp.nlex + sp.nrex, sz))
            sats.append(sp)
        nsmall = sum(s.size <= self.nsm for s in sats)
        sats.sort(key=SatPos.sort_key)
        out = []
        added = 0
        j = 0
        while j < nsmall and added < maxelt:
            s = sats[j]
            r = Random1toN()
            r.init(s.size, self.all)
            out.append((s, None, r))
            added += s.size
            j += 1
        if added >= maxelt or nsmall == len(sats):
            return out, added
        sampler = RowSampler(sats[nsmall:])
        rands2 = [Random1toN() for _ in sats]
        while added < maxelt and added < nelt:
            ri = sampler.next(rnd) + nsmall
            if not rands2[ri].inited():
                rands2[ri].init(sats[ri].size, self.all)
            r = rands2[ri].next(rnd)
            if rands2[ri].done():
                sampler.finished(ri - nsmall)
            src = sats[ri]
            s = SatPos(src.topf + r, 0, 1, src.key_len, src.fw, src.offidx, src.rdoff, src.seedlen, src.nlex, src.nrex, src.orig_size)
            one = Random1toN()
            one.init(1, self.all)
            out.append((s, None, one))
            added += 1
        return out, added

    # --------------------------------------------------------------------------------------- extendSeeds
    def extend_seeds(self, sh, ee_exact):
        sc, rnd, rdlen = self.sc, self.rnd, self.cur.rdlen
        nonz = sh["nonz"] if sh else 0
        ee_mode = bool(ee_exact or self.cur.mm1)
        first_ee = first_extend = True
        n_ug_fail = n_dp_fail = 0
        nelt_left = 0
        satpos = []
        while True:
            if ee_mode:
                if first_ee:
                    first_ee = False
                    satpos, _ = self._ee_sa_tups(ee_exact, self.max_iters)
                else:
                    ee_mode = False
            if not ee_mode:
                if nonz == 0:
                    return EXHAUSTED
                if self.cur.minsc == self.cur.perfect:
                    return PERFECT
                if first_extend:
                    satpos, nelt = yield from self._prioritize(sh, self.max_iters)
                    nelt_left = nelt
                    first_extend = False
                if nelt_left == 0:
                    break
            for sp, eehit, rands in satpos:
                if ee_mode and eehit.score < self.cur.minsc:
                    return PERFECT
                is_small = sp.size < self.nsm
                fw = sp.fw
                rdoff = sp.rdoff
                if not fw:
                    rdoff = rdlen - rdoff - sp.seedlen
                first = True
                while (not rands.done()) and (first or is_small or ee_mode):
                    if self.cur.minsc == self.cur.perfect:
                        if not ee_mode or eehit.score < self.cur.perfect:
                            return PERFECT
                    elif ee_mode and eehit.score < self.cur.minsc:
                        break
                    if self.n_dps >= self.max_dp or self.n_ugs >= self.max_ug or self.n_iters >= self.max_iters:
                        return HARD_LIMIT
                    self.n_iters += 1
                    first = False
                    elt = rands.next(rnd)
                    joined = yield ("resolve", (sp.topf + elt,))
                    if not ee_mode:
                        nelt_left -= 1
                    ok, tidx, toff, tlen, straddled = yield ("joined_to_text", (sp.key_len, joined, ee_mode,))
                    if not ok:
                        continue
                    refoff = toff - rdoff
                    if self.cur.seen.present(tidx, fw, refoff):
                        self.n_red += 1
                        continue
                    read_gaps = ref_gaps = 0
                    ungapped = False
                    if not ee_mode:
                        read_gaps = sc.max_read_gaps(self.cur.minsc, rdlen)
                        ref_gaps = sc.max_ref_gaps(self.cur.minsc, rdlen)
                        ungapped = read_gaps == 0 and ref_gaps == 0
                    found_alns = None                 # list of Aln for EE / ungapped, or a DP attempt iterator
                    state = 0
                    if ee_mode:
                        ed = [] if eehit.edit is None else [(eehit.edit[0], eehit.edit[1], eehit.edit[2], 3)]
                        a = Aln(tidx, refoff, fw, eehit.score, rdlen, ed, eehit.ns(), eehit.refns(), True)
                        found_alns = [a]
                        state = 1
                        self.cur.seen.add(tidx, fw, refoff, 1)
                    elif ungapped:
                        rc, a = yield ("ungapped", (self.cur.codes, self.cur.quals, fw, tidx, refoff, tlen, self.cur.minsc,))
                        self.cur.seen.add(tidx, fw, refoff, 1)
                        self.n_ugs += 1
                        if rc == 0:
                            n_ug_fail += 1
                            if n_ug_fail >= self.streak:
                                return SOFT_LIMIT
                            continue
                        elif rc == -1:
                            n_ug_fail += 1
                            if n_ug_fail >= self.streak:
                                return SOFT_LIMIT
                        else:
                            n_ug_fail = 0
                            found_alns = [a]
                            state = 2
                    dp = None
                    if state == 0:
                        found, rect = policy.frame_seed_extension_rect(refoff, rdlen, tlen, read_gaps, ref_gaps, self.cur.nceil, self.maxhalf)
                        self.cur.seen.add(tidx, fw, refoff, 1)
                        if not found:
                            continue
                        self.cur.seen.add(tidx, fw, rect.refl_pretrim + rect.corel, rect.corer - rect.corel + 1)
                        dp = yield ("dp", (self.cur.codes, self.cur.quals, fw, tidx, rect, self.cur.minsc, sc.n_ceil_raw(rdlen),))
                        self.n_dps += 1
                        if not dp["found"]:
                            n_dp_fail += 1
                            if n_dp_fail >= self.streak:
                                return SOFT_LIMIT
                            continue
                        n_dp_fail = 0
                        dp["cursor"] = 0
                        dp["u8"] = self._dp_u8(dp, self.cur.minsc, self.cur.quals)
                    first_inner = True
                    while True:
                        if state != 0:
                            if not first_inner:
                                break
                            a = found_alns[0]
                        else:
                            a = yield from self._next_alignment(dp, tidx, self.cur.minsc, rdlen)
                            if a is None:
                                break
                        first_inner = False
                        # (alignments falling off the reference are clipped only with --local / overhangs enabled)
                        if self.red.overlap(a):
                            continue
                        self.red.add(a)
                        if self.sink.report(a):
                            return FULFILLED
                        if self.tighten > 0 and self.mmode and self.sink.best2 != MIN_I64:
                            best, best2 = self.sink.best, self.sink.best2
                            if self.tighten == 1:
                                if best >= self.cur.minsc:
                                    self.cur.minsc = best
                                    if self.cur.minsc < self.cur.perfect and best == best2:
                                        self.cur.minsc += 1
                            elif self.tighten == 2:
                                if best2 >= self.cur.minsc:
                                    self.cur.minsc = best2
                                    if self.cur.minsc < self.cur.perfect:
                                        self.cur.minsc += 1
                            else:
                                diff = best - best2
                                bot = best2 + (diff * 3) // 4              # diff >= 0
                                if bot >= self.cur.minsc:
                                    self.cur.minsc = bot
                                    if self.cur.minsc < self.cur.perfect:
                                        self.cur.minsc += 1
        return EXHAUSTED

    def _dp_u8(self, dp, minsc, quals):
        """did SwAligner::align stay on the 8-bit matrices (aligner_sw.cpp:514-600)?  End-to-end: the minimum score must
        fit; local: no cell may reach 255 - bias, bias = the largest penalty of the query profile (aligner_swsse_loc_u8.cpp:97-110)"""
        if not self.local:
            return minsc >= -254
        bias = max([self.sc.mm_penalty(int(q) - 33) for q in quals] + [self.sc.n_pen])
        return dp["best"] + bias < 255

    def _next_alignment(self, dp, tidx, minsc, rdlen):
        """SwAligner::nextAlignment (aligner_sw.cpp:737-1146) over the backend's attempt list: candidates below the
        current minimum score are skipped without touching the RNG, every backtrace attempt reseeds it."""
        rnd = self.rnd
        att = dp["attempts"]
        while dp["cursor"] < len(att):
            cand_score, ai = att[dp["cursor"]]
            dp["cursor"] += 1
            if cand_score < minsc:
                continue
            reseed = (rnd.next_u32() + 1) & 0xffffffff
            rnd.init((reseed + 1) & 0xffffffff if dp["u8"] else reseed)
            if ai >= 0:
                al = dp["alns"][ai]
                ed = [tuple(e) for e in al["edits"]]
                # AlnRes::refNs: ambiguous reference characters under the alignment (XN:i)
                a = Aln(tidx, al["refoff"], bool(al["fw"]), al["score"], rdlen, ed, al["ns"], 0, False, al["trim5"], al["trim3"])
                a.refns = yield ("count_ref_ns", (tidx, a.refoff, a.ref_extent,))
                return a
        return None

    # ---------------------------------------------------------------------------------------- finishRead
    def finish_read(self, res: ReadResult) -> ReadResult:
        rnd = self.rnd
        alns = self.sink.alns
        res.n_alns = len(alns)
        res.maxed = self.sink.exit_m
        res.counters = dict(ZI=self.n_iters, XD=self.n_dps, XU=self.n_ugs, YR=self.n_red)
        if not alns:
            return res
        # selectByScore (aln_sink.cpp:1477-1628): descending by score, index descending within ties, tie streaks shuffled
        buf = sorted(((a.score, i) for i, a in enumerate(alns)), reverse=True)
        shuffle_equal_streaks(buf, lambda t: t[0], rnd)
        best = alns[buf[0][1]]
        res.aligned, res.aln = True, best
        res.xs = buf[1][0] if len(buf) > 1 else None
        res.mapq = self._mapq(best.score, res.xs, self.sc.min_score(self.cur.rdlen), self.cur.perfect)
        # ReportingState::getReport: -k short circuit -> khits alignments, else min(found, khits)
        num = self.khits if self.sink.exit_k else min(len(alns), self.khits)
        res.secondary = [alns[buf[i][1]] for i in range(1, min(num, len(buf)))]
        return res

    def _mapq(self, best, secbest, sc_min, perfect):
        """BowtieMapq2::mapq incl. its 255 case: without -M (no "canMax") and no second-best score the search says nothing
        about uniqueness (unique.h:201-205; the search is never flagged exhaustive)"""
        if not self.mmode and secbest is None:
            return 255
        return policy.mapq_v2(best, secbest, sc_min, perfect, not self.local)


def aln_to_ops(a: Aln, codes):
    """op string of include/bt2g.h (last aligned read row first; every op that consumes a reference base carries its
    code) from an alignment in the reference's Edit representation."""
    from .lib import OP_MATCH, OP_MM, OP_READGAP, OP_REFGAP
    code = {ord(c): i for i, c in enumerate("ACGTN")}
    seq = codes if a.fw else np.array([4 if c > 3 else 3 - c for c in codes[::-1]], dtype=np.uint8)
    ed = edits_left_to_right(a)
    fwd, k = [], 0
    row0 = a.trim_left
    for rel in range(a.ext):
        while k < len(ed) and ed[k][0] == rel and ed[k][3] == 1:
            fwd.append(OP_READGAP | (code[ed[k][1]] << 2))
            k += 1
        if k < len(ed) and ed[k][0] == rel:
            fwd.append(OP_REFGAP if ed[k][3] == 2 else (OP_MM | (code[ed[k][1]] << 2)))
            k += 1
        else:
            fwd.append(OP_MATCH | (int(seq[row0 + rel]) << 2))
    assert k == len(ed)
    return fwd[::-1]


# =====================================================================================================================
# Paired-end reads: multiseedSearchWorker's two-mate flow (bt2_search.cpp:3400-4250), SwDriver::extendSeedsPaired
# (aligner_sw_driver.cpp:1582-2615), ReportingState / AlnSinkWrap for pairs (aln_sink.cpp:26-330, 643-1070).
# Program defaults: --fr, -I 0 -X 500, discordant and mixed (unpaired) alignments reported.

class PairedSink:
    def __init__(self, khits=1, mhits=50, discord=True, mixed=True, mmode=True):
        self.khits, self.mhits, self.mmode = khits, mhits, mmode
        self.exit_concord_k = False
        self.exit_unp_k = [False, False]
        self.rs1, self.rs2, self.rs1u, self.rs2u = [], [], [], []
        self.done_concord = False
        self.done_discord = not discord
        self.done_unp = [not mixed, not mixed]
        self.exit_concord_m = False
        self.exit_unp_m = [False, False]
        self.discord_exit_noaln = False
        self.nconcord = 0
        self.nunp = [0, 0]
        self.done = False
        self.best_pair = self.best2_pair = MIN_I64
        self.best_unp = [MIN_I64, MIN_I64]
        self.best2_unp = [MIN_I64, MIN_I64]

    def _update_done(self):
        self.done = self.done_unp[0] and self.done_unp[1] and self.done_discord and self.done_concord

    def report(self, a1, a2):
        """AlnSinkWrap::report: a pair when both are given, else an unpaired alignment of the given mate"""
        if a1 is not None and a2 is not None:
            self.nconcord += 1
            if not self.mmode and self.nconcord >= self.khits:
                self.done_concord, self.exit_concord_k = True, True
            elif self.mmode and self.nconcord > self.mhits:
                self.done_concord, self.exit_concord_m = True, True
            self.done_discord = True
            if self.done_concord and not self.exit_concord_m:
                # closed by -k: the unpaired categories are trumped (a category closed by -M does not trump them)
                self.done_unp = [True, True]
            self._update_done()
            self.rs1.append(a1)
            self.rs2.append(a2)
            score = a1.score + a2.score
            if score > self.best_pair:
                self.best2_pair, self.best_pair = self.best_pair, score
            elif score > self.best2_pair:
                self.best2_pair = score
        else:
            m = 0 if a1 is not None else 1
            a = a1 if a1 is not None else a2
            self.nunp[m] += 1
            if not self.done_unp[m]:
                if not self.mmode and self.nunp[m] >= self.khits:
                    self.done_unp[m], self.exit_unp_k[m] = True, True
                    self._update_done()
                elif self.mmode and self.nunp[m] > self.mhits:
                    self.done_unp[m], self.exit_unp_m[m] = True, True
                    self._update_done()
            if self.nunp[m] > 1:
                self.done_discord = True
            (self.rs1u if m == 0 else self.rs2u).append(a)
            if a.score > self.best_unp[m]:
                self.best2_unp[m], self.best_unp[m] = self.best_unp[m], a.score
            elif a.score > self.best2_unp[m]:
                self.best2_unp[m] = a.score
        return self.done

    def done_with_mate(self, mate1):
        m = 0 if mate1 else 1
        if not self.done_unp[m] or not self.done_concord:
            return False
        if not self.done_discord and self.nunp[m] == 0:
            return False
        return True

    def done_unpaired(self, mate1):
        return self.done_unp[0 if mate1 else 1]


@dataclass
class PairResult:
    pair_type: int = 0             # include/bt2g.h: 0 none, 1 concordant, 2 both aligned (discordant or unpaired), 3 one mate
    mates: list = None             # two ReadResult
    counters: dict = None
    n_concord: int = 0
    secondary_pairs: list = None   # -k / -a: the further selected concordant pairs, in report order


class PairedPolicyEngine(PolicyEngine):
    def __init__(self, backend, preset="sensitive", seed=0, sc=None, pe=None, local=False, discord=True, mixed=True, **kw):
        """pe = PairedEndPolicy (-I / -X / --fr --rf --ff / --dovetail / --no-contain / --no-overlap);
        discord / mixed = not --no-discordant / not --no-mixed"""
        super().__init__(backend, preset, seed, sc, local, **kw)
        self.pe = pe or policy.PairedEndPolicy(local=local)
        self.discord, self.mixed = discord, mixed

    def align_pair(self, codes1, quals1, name1, codes2, quals2, name2) -> PairResult:
        return self._drive(self.pair_steps(codes1, quals1, name1, codes2, quals2, name2))

    def pair_steps(self, codes1, quals1, name1, codes2, quals2, name2):
        sc = self.sc
        m = []
        for codes, quals, name in ((codes1, quals1, name1), (codes2, quals2, name2)):
            codes = np.asarray(codes, dtype=np.uint8)
            quals = np.asarray(quals, dtype=np.uint8)
            rdlen = len(codes)
            c = MateCtx(codes, quals, name, rdlen, sc.min_score(rdlen) if rdlen else 0, sc.perfect_score(rdlen), sc.n_ceil(rdlen) if rdlen else 0)
            if rdlen < 2:
                c.filt, c.filtered = False, "LN"
            elif int((codes > 3).sum()) > sc.n_ceil(rdlen):
                c.filt, c.filtered = False, "NS"
            elif sc.perfect_score(rdlen) < sc.min_score(rdlen):
                c.filt, c.filtered = False, "SC"
            m.append(c)
        self.m = m
        both = m[0].filt and m[1].filt
        s1 = policy.gen_rand_seed(m[0].codes, m[0].quals, name1, self.seed)
        s2 = policy.gen_rand_seed(m[1].codes, m[1].quals, name2, self.seed)
        rnd = self.rnd = RandomSource((s1 ^ s2) if both else s1)
        interval = [policy.seed_interval(self.pre.ival, c.rdlen, both) if c.rdlen else 1 for c in m]
        streak = self.streak                             # -D, raised by -k / unbounded with -a (set in __init__)
        nrounds_all = self.n_seed_rounds
        if both:
            streak = -(-streak // 2)
            nrounds_all = -(-nrounds_all // 2)
        self.streak_cur = streak
        self.red = RedundantAlns()
        self.red_mate = [RedundantAlns(), RedundantAlns()]
        self.sink = PairedSink(self.khits, self.mhits, self.discord, self.mixed, self.mmode)
        self.n_iters = self.n_dps = self.n_ugs = self.n_red = self.n_mate_dps = 0
        # bt2_search.cpp:3419-3426: --nofw / --norc refer to the fragment; which strand of a mate that is depends on --fr/--rf/--ff
        m1fw = self.pe.pol in (policy.PE_POLICY_FF, policy.PE_POLICY_FR)
        m2fw = self.pe.pol in (policy.PE_POLICY_FF, policy.PE_POLICY_RF)
        nofw = [self.gnofw if m1fw else self.gnorc, self.gnofw if m2fw else self.gnorc]
        norc = [self.gnorc if m1fw else self.gnofw, self.gnorc if m2fw else self.gnofw]
        done = [not m[0].filt, not m[1].filt]
        sink = self.sink
        matemap = [0, 1]
        nelt = [0, 0]
        mined = [[0, 0], [0, 0]]

        def after(ret, mate):
            if ret == FULFILLED:
                if sink.done_with_mate(mate == 0):
                    done[mate] = True
                if sink.done_with_mate(mate == 1):
                    done[mate ^ 1] = True
            elif ret in (PERFECT, HARD_LIMIT):
                done[mate] = True

        # ---- exact end-to-end
        for mate in matemap:
            c = m[mate]
            if not c.filt or done[mate] or sink.done_with_mate(mate == 0):
                continue
            ne, mi, tb = yield ("exact_sweep", (c.codes, nofw[mate], norc[mate],))
            nelt[mate] = ne
            mined[mate] = [int(mi[0]), int(mi[1])]
            c.ee = []
            if tb[1] > tb[0]:
                c.ee.append(EEHit(int(tb[0]), int(tb[1]), True, c.perfect))
            if tb[3] > tb[2]:
                c.ee.append(EEHit(int(tb[2]), int(tb[3]), False, c.perfect))
        matemap = [1, 0] if (nelt[0] > 0 and nelt[1] > 0 and nelt[0] > nelt[1]) else [0, 1]
        for mate in matemap:
            c = m[mate]
            if nelt[mate] == 0:
                c.ee = []
                continue
            if sink.done_with_mate(mate == 0):
                c.ee = []
                done[mate] = True
                continue
            ret = yield from self.extend_seeds_paired(mate, None, c.ee)
            c.ee = []
            after(ret, mate)
            if not done[mate] and c.minsc == c.perfect:
                done[mate] = True
        # ---- 1-mismatch end-to-end
        for mate in matemap:
            c = m[mate]
            if not c.filt or done[mate]:
                c.mm1 = []
                nelt[mate] = 0
                continue
            nelt[mate] = 0
            yfw, yrc = mined[mate][0] <= 1 and not nofw[mate], mined[mate][1] <= 1 and not norc[mate]
            if yfw or yrc:
                hits = yield ("one_mm", (c.codes, c.quals, c.minsc, not yfw, not yrc,))
                c.mm1 = [EEHit(int(h[0]), int(h[1]), bool(h[6]), int(h[5]), (int(h[2]), int(h[3]), int(h[4]))) for h in hits]
                nelt[mate] = sum(h.bot - h.top for h in c.mm1)
        matemap = [1, 0] if (nelt[0] > 0 and nelt[1] > 0 and nelt[0] > nelt[1]) else [0, 1]
        for mate in matemap:
            c = m[mate]
            if nelt[mate] == 0:
                continue
            if sink.done_with_mate(mate == 0):
                done[mate] = True
                continue
            ret = yield from self.extend_seeds_paired(mate, None, [])
            c.mm1 = []
            after(ret, mate)
            if not done[mate] and c.minsc == c.perfect:
                done[mate] = True
        # ---- seed rounds
        nrounds = [min(nrounds_all, interval[0]), min(nrounds_all, interval[1])]
        L = self.pre.seed_len
        for roundi in range(self.n_seed_rounds):
            for c in m:
                c.sh = None
            for mate in matemap:
                c = m[mate]
                if done[mate] or sink.done_with_mate(mate == 0):
                    done[mate] = True
                    continue
                if roundi >= nrounds[mate] or interval[mate] <= roundi:
                    continue
                offset = (interval[mate] * roundi) // nrounds[mate]
                if offset > 0 and min(L, c.rdlen) + offset > c.rdlen:
                    continue
                hits = yield ("seed_search", (c.codes, c.quals, min(L, c.rdlen), interval[mate], offset, nofw[mate], norc[mate],))
                nfw = [max(0, int(h[1]) - int(h[0])) for h in hits[0]]
                nrc = [max(0, int(h[1]) - int(h[0])) for h in hits[1]]
                nonz = sum(x > 0 for x in nfw) + sum(x > 0 for x in nrc)
                if nonz == 0:
                    done[mate] = True
                    break
                c.sh = dict(hits=hits, interval=interval[mate], offset=offset, seedlen=min(L, c.rdlen), nonz=nonz,
                            nelt=sum(nfw) + sum(nrc), nfw=nfw, nrc=nrc)
            uniq = [0.0, 0.0]
            for i, c in enumerate(m):
                if c.sh:
                    uniq[i] = sum(1.0 / float(x * x) for x in c.sh["nfw"] + c.sh["nrc"] if x > 0)
            matemap = [1, 0] if (m[0].sh and m[1].sh and uniq[1] > uniq[0]) else [0, 1]
            for mate in matemap:
                c = m[mate]
                if done[mate] or sink.done_with_mate(mate == 0):
                    done[mate] = True
                    continue
                if not c.sh:
                    continue
                c.sh["ranks"] = policy.rank_seed_hits(c.sh["nfw"], c.sh["nrc"], rnd, self.all)
                ret = yield from self.extend_seeds_paired(mate, c.sh, [])
                after(ret, mate)
            for mate in (0, 1):
                c = m[mate]
                if not done[mate] and c.sh and c.sh["nelt"] // c.sh["nonz"] < self.seed_boost_thresh:
                    done[mate] = True
        return self.finish_pair()

    # ------------------------------------------------------------------------------------ extendSeedsPaired
    def _tightened_pair_score(self, best_pair_score):
        sink = self.sink
        if self.tighten == 1:
            ps = sink.best_pair
        elif self.tighten == 2:
            ps = sink.best2_pair
        else:
            ps = sink.best2_pair + ((sink.best_pair - sink.best2_pair) * 3) // 4
        if self.tighten == 1 and ps < best_pair_score and sink.best_pair == sink.best2_pair:
            ps += 1
        if self.tighten >= 2 and ps < best_pair_score:
            ps += 1
        return ps

    def extend_seeds_paired(self, ai, sh, ee_exact):
        sc, rnd, sink, pe = self.sc, self.rnd, self.sink, self.pe
        anchor1 = ai == 0
        c = self.cur = self.m[ai]
        o = self.m[ai ^ 1]
        rdlen, ordlen = c.rdlen, o.rdlen
        opp_filt = not o.filt
        operfect = o.perfect
        best_pair_score = c.perfect + operfect
        if self.tighten > 0 and self.mmode and sink.best2_pair != MIN_I64:
            nc = self._tightened_pair_score(best_pair_score) - operfect
            if nc > c.minsc:
                c.minsc = nc
        nonz = sh["nonz"] if sh else 0
        ee_mode = bool(ee_exact or c.mm1)
        first_ee = first_extend = True
        n_ee_fail = n_ug_fail = n_dp_fail = 0
        nelt_left = 0
        satpos = []
        mate_streaks = []
        sw_mate_immediately = True
        streak = self.streak_cur
        while True:
            if ee_mode:
                if first_ee:
                    first_ee = False
                    satpos, _ = self._ee_sa_tups(ee_exact, self.max_iters)
                    mate_streaks = [0] * len(satpos)
                else:
                    ee_mode = False
            if not ee_mode:
                if nonz == 0:
                    return EXHAUSTED
                if self.mmode and c.minsc == c.perfect:
                    return PERFECT
                if first_extend:
                    satpos, nelt = yield from self._prioritize(sh, self.max_iters)
                    nelt_left = nelt
                    first_extend = False
                    mate_streaks = [0] * len(satpos)
                if nelt_left == 0:
                    break
            for si, (sp, eehit, rands) in enumerate(satpos):
                if ee_mode and eehit.score < c.minsc:
                    return PERFECT
                is_small = sp.size < self.nsm
                fw = sp.fw
                rdoff = sp.rdoff
                if not fw:
                    rdoff = rdlen - rdoff - sp.seedlen
                first = True
                while (not rands.done()) and (first or is_small or ee_mode):
                    if c.minsc == c.perfect:
                        if not ee_mode or eehit.score < c.perfect:
                            return PERFECT
                    elif ee_mode and eehit.score < c.minsc:
                        break
                    if self.n_dps >= self.max_dp or self.n_mate_dps >= self.max_dp or self.n_ugs >= self.max_ug or self.n_iters >= self.max_iters:
                        return HARD_LIMIT
                    if ee_mode and n_ee_fail >= streak:
                        return SOFT_LIMIT
                    if not ee_mode and (n_dp_fail >= streak or n_ug_fail >= streak):
                        return SOFT_LIMIT
                    if mate_streaks[si] >= self.max_mate_streak:
                        rands.cur = rands.n                    # Random1toN::setDone
                        break
                    self.n_iters += 1
                    first = False
                    elt = rands.next(rnd)
                    joined = yield ("resolve", (sp.topf + elt,))
                    nelt_left -= 1
                    ok, tidx, toff, tlen, straddled = yield ("joined_to_text", (sp.key_len, joined, ee_mode,))
                    if not ok:
                        continue
                    refoff = toff - rdoff
                    if c.seen.present(tidx, fw, refoff):
                        self.n_red += 1
                        continue
                    read_gaps = ref_gaps = 0
                    ungapped = False
                    if not ee_mode:
                        read_gaps = sc.max_read_gaps(c.minsc, rdlen)
                        ref_gaps = sc.max_ref_gaps(c.minsc, rdlen)
                        ungapped = read_gaps == 0 and ref_gaps == 0
                    state = 0
                    fixed = None
                    if ee_mode:
                        ed = [] if eehit.edit is None else [(eehit.edit[0], eehit.edit[1], eehit.edit[2], 3)]
                        fixed = Aln(tidx, refoff, fw, eehit.score, rdlen, ed, eehit.ns(), eehit.refns(), True)
                        state = 1
                        c.seen.add(tidx, fw, refoff, 1)
                        n_ee_fail += 1
                    elif ungapped:
                        rc, a = yield ("ungapped", (c.codes, c.quals, fw, tidx, refoff, tlen, c.minsc,))
                        c.seen.add(tidx, fw, refoff, 1)
                        self.n_ugs += 1
                        n_ug_fail += 1
                        if rc == 0:
                            continue
                        if rc == 1:
                            fixed = a
                            state = 2
                    dp = None
                    if state == 0:
                        found, rect = policy.frame_seed_extension_rect(refoff, rdlen, tlen, read_gaps, ref_gaps, c.nceil, self.maxhalf)
                        c.seen.add(tidx, fw, refoff, 1)
                        if not found:
                            continue
                        c.seen.add(tidx, fw, rect.refl_pretrim + rect.corel, rect.corer - rect.corel + 1)
                        dp = yield ("dp", (c.codes, c.quals, fw, tidx, rect, c.minsc, sc.n_ceil_raw(rdlen),))
                        self.n_dps += 1
                        n_dp_fail += 1
                        if not dp["found"]:
                            continue
                        dp["cursor"] = 0
                        dp["u8"] = self._dp_u8(dp, c.minsc, c.quals)
                    first_inner = True
                    found_concordant = False
                    while True:
                        if state != 0:
                            if not first_inner:
                                break
                            a = fixed
                        else:
                            a = yield from self._next_alignment(dp, tidx, c.minsc, rdlen)
                            if a is None:
                                break
                        first_inner = False
                        if self.red.overlap(a):
                            continue
                        self.red.add(a)
                        if sink.done_with_mate(not anchor1) and not sink.done_with_mate(anchor1):
                            sw_mate_immediately = False
                        if sw_mate_immediately:
                            found_mate = not opp_filt
                            ominsc_cur = o.minsc
                            odp = None
                            if found_mate:
                                if self.tighten > 0 and self.mmode and sink.best2_pair != MIN_I64:
                                    nc = self._tightened_pair_score(best_pair_score) - a.score
                                    if nc > ominsc_cur:
                                        ominsc_cur = nc
                                ordgaps = sc.max_read_gaps(ominsc_cur, ordlen)
                                orfgaps = sc.max_ref_gaps(ominsc_cur, ordlen)
                                om = pe.other_mate(anchor1, fw, a.refoff, ordlen + ordgaps, tlen,
                                                   rdlen if anchor1 else ordlen, ordlen if anchor1 else rdlen)
                                found_mate = om is not None
                            if found_mate:
                                oleft, oll, olr, orl, orr, ofw = om
                                found_mate, orect = policy.frame_find_mate_rect(not oleft, oll, olr, orl, orr, ordlen, tlen, ordgaps, orfgaps,
                                                                                o.nceil, self.maxhalf)
                            if found_mate:
                                odp = yield ("dp", (o.codes, o.quals, ofw, tidx, orect, ominsc_cur, sc.n_ceil_raw(ordlen),))
                                self.n_mate_dps += 1
                                found_mate = bool(odp["found"])
                                if found_mate:
                                    odp["cursor"] = 0
                                    odp["u8"] = self._dp_u8(odp, ominsc_cur, o.quals)
                            did_anchor = False
                            brk = False
                            while True:
                                oa = None
                                if found_mate:
                                    oa = yield from self._next_alignment(odp, tidx, ominsc_cur, ordlen)
                                    found_mate = oa is not None
                                if found_mate:
                                    if not self.red.overlap(oa):
                                        self.red.add(oa)
                                    oext = oa.ref_extent
                                    if oa.refoff < 0 or oa.refoff + oext > tlen:
                                        found_mate = False          # falls off the reference (no overhangs)
                                pair_cl = policy.PE_ALS_DISCORD
                                if found_mate:
                                    aext = a.ref_extent
                                    a1, a2 = (a, oa) if anchor1 else (oa, a)
                                    l1, l2 = (aext, oext) if anchor1 else (oext, aext)
                                    pair_cl = pe.classify_pair(a1.refoff, l1, a1.fw, a2.refoff, l2, a2.fw)
                                if sink.done_concord:
                                    found_mate = False
                                if found_mate:
                                    done_unpaired = False
                                    if not anchor1 or not did_anchor:
                                        if anchor1:
                                            did_anchor = True
                                        r1 = a if anchor1 else oa
                                        if not self.red_mate[0].overlap(r1):
                                            self.red_mate[0].add(r1)
                                            if sink.report(r1, None):
                                                done_unpaired = True
                                    if anchor1 or not did_anchor:
                                        if not anchor1:
                                            did_anchor = True
                                        r2 = oa if anchor1 else a
                                        if not self.red_mate[1].overlap(r2):
                                            self.red_mate[1].add(r2)
                                            if sink.report(None, r2):
                                                done_unpaired = True
                                    done_paired = False
                                    if pair_cl != policy.PE_ALS_DISCORD:
                                        found_concordant = True
                                        if sink.report(a if anchor1 else oa, oa if anchor1 else a):
                                            done_paired = True
                                        elif self.tighten > 0 and self.mmode and sink.best2_pair != MIN_I64:
                                            nc = self._tightened_pair_score(best_pair_score) - operfect
                                            if nc > c.minsc:
                                                c.minsc = nc
                                                if c.minsc > a.score:
                                                    brk = True
                                    if brk:
                                        break
                                    if done_paired or done_unpaired:
                                        return FULFILLED
                                    if sink.done_with_mate(anchor1):
                                        return FULFILLED
                                elif (self.mixed or self.discord) and not did_anchor:
                                    did_anchor = True
                                    if not sink.done_unpaired(anchor1):
                                        red = self.red_mate[0 if anchor1 else 1]
                                        if not red.overlap(a):
                                            red.add(a)
                                            if sink.report(a if anchor1 else None, None if anchor1 else a):
                                                return FULFILLED
                                    if sink.done_with_mate(anchor1):
                                        return FULFILLED
                                if oa is None:
                                    break
                        elif self.mixed or self.discord:
                            if not sink.done_unpaired(anchor1):
                                red = self.red_mate[0 if anchor1 else 1]
                                if not red.overlap(a):
                                    red.add(a)
                                    if sink.report(a if anchor1 else None, None if anchor1 else a):
                                        return FULFILLED
                            if sink.done_with_mate(anchor1):
                                return FULFILLED
                    if found_concordant:
                        mate_streaks[si] = 0
                        if state == 2:
                            n_ug_fail = 0
                        elif state == 1:
                            n_ee_fail = 0
                        else:
                            n_dp_fail = 0
                    else:
                        mate_streaks[si] += 1
        return EXHAUSTED

    # ------------------------------------------------------------------------------------------ finishRead
    def _select(self, rs1, rs2, rs1u, rs2u):
        """selectByScore for a list of pairs (rs2 given) or of unpaired alignments: -> (index, best-unchosen info)"""
        rnd = self.rnd
        buf = sorted(((a.score + (rs2[i].score if rs2 is not None else 0), i) for i, a in enumerate(rs1)), reverse=True)
        shuffle_equal_streaks(buf, lambda t: t[0], rnd)
        sel = buf[0][1]
        out = dict(sel=sel, order=[t[1] for t in buf], unchosen_u=None, unchosen_p=[None, None], unchosen_c=None)
        if rs2 is not None:
            for k, (rsu, chosen) in enumerate(((rs1u, rs1[sel]), (rs2u, rs2[sel]))):
                best = None
                for a in rsu:
                    if (a.tidx, a.refoff, a.fw) == (chosen.tidx, chosen.refoff, chosen.fw):
                        continue
                    if best is None or a.score > best:
                        best = a.score
                out["unchosen_p"][k] = best
            if len(buf) > 1:
                out["unchosen_c"] = buf[1][0]
        elif len(buf) > 1:
            out["unchosen_u"] = rs1[buf[1][1]].score
        return out

    def finish_pair(self) -> PairResult:
        sink, m, sc = self.sink, self.m, self.sc
        res = PairResult(mates=[ReadResult(filtered=m[0].filtered), ReadResult(filtered=m[1].filtered)])
        res.counters = dict(ZI=self.n_iters, XD=self.n_dps, XU=self.n_ugs, YR=self.n_red)
        res.n_concord = sink.nconcord
        mn = [sc.min_score(c.rdlen) if c.rdlen else 0 for c in m]
        # ReportingState::finish + getReport (aln_sink.cpp:131-260)
        if sink.nconcord > 0:
            s = self._select(sink.rs1, sink.rs2, sink.rs1u, sink.rs2u)
            a1, a2 = sink.rs1[s["sel"]], sink.rs2[s["sel"]]
            mq = self._mapq(a1.score + a2.score, s["unchosen_c"], mn[0] + mn[1], m[0].perfect + m[1].perfect)
            for k, a in enumerate((a1, a2)):
                r = res.mates[k]
                r.aligned, r.aln, r.xs, r.mapq = True, a, s["unchosen_p"][k], mq
            res.pair_type = 1
            num = self.khits if sink.exit_concord_k else min(sink.nconcord, self.khits)
            res.secondary_pairs = [(sink.rs1[j], sink.rs2[j]) for j in s["order"][1:num]]
            return res
        discord = (not sink.done_discord) and sink.nunp[0] == 1 and sink.nunp[1] == 1
        if discord:
            # prepareDiscordants + selectByScore over the single pair
            s = self._select([sink.rs1u[0]], [sink.rs2u[0]], sink.rs1u, sink.rs2u)
            a1, a2 = sink.rs1u[0], sink.rs2u[0]
            mq = self._mapq(a1.score + a2.score, None, mn[0] + mn[1], m[0].perfect + m[1].perfect)
            for k, a in enumerate((a1, a2)):
                r = res.mates[k]
                r.aligned, r.aln, r.xs, r.mapq = True, a, None, mq
            res.pair_type = 2
            return res
        for k, rsu in enumerate((sink.rs1u, sink.rs2u)):
            if not rsu or not self.mixed:              # --no-mixed: ReportingState::getReport returns before the unpaired counts
                continue
            s = self._select(rsu, None, None, None)
            a = rsu[s["sel"]]
            r = res.mates[k]
            r.aligned, r.aln, r.xs = True, a, s["unchosen_u"]
            r.mapq = self._mapq(a.score, s["unchosen_u"], mn[k], m[k].perfect)
            r.n_alns = len(rsu)
            num = self.khits if sink.exit_unp_k[k] else min(len(rsu), self.khits)
            r.secondary = [rsu[j] for j in s["order"][1:num]]
        n_al = sum(r.aligned for r in res.mates)
        res.pair_type = 2 if n_al == 2 else (3 if n_al == 1 else 0)
        return res
