"""CPU: host policy arithmetic vs the reference (oracle/_ref glue)."""
import ctypes as C

import numpy as np
import pytest

from bowtie2_b200 import policy
from oracle_lib import Reference, have_reference

i64 = C.c_int64


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_scoring_and_framer_match_reference(lambda_index):
    R = Reference(lambda_index)
    L = R.lib
    L.ref_frame_seed_rect.argtypes = [i64, C.c_uint64, i64, C.c_uint64, C.c_uint64, i64, C.c_uint64, C.POINTER(i64)]
    L.ref_score_params.argtypes = [C.c_void_p, C.c_int, i64, C.c_uint64, C.POINTER(i64)]
    rng = np.random.default_rng(0)
    for local in (False, True):
        sc = policy.Scoring.default(local)
        for rdlen in list(range(20, 60)) + [75, 100, 150, 250, 300, 500]:
            minsc = sc.min_score(rdlen)
            for ms in {minsc, minsc + 7, min(minsc + 30, sc.perfect_score(rdlen))}:
                out = (i64 * 4)()
                L.ref_score_params(R.h, int(local), ms, rdlen, out)
                assert (sc.max_read_gaps(ms, rdlen), sc.max_ref_gaps(ms, rdlen), sc.perfect_score(rdlen),
                        sc.n_ceil_raw(rdlen)) == tuple(out), (local, rdlen, ms)
            rg, fg = sc.max_read_gaps(minsc, rdlen), sc.max_ref_gaps(minsc, rdlen)
            for off in list(rng.integers(-100, 48600, 30)) + [-70, -1, 0, 48502 - rdlen, 48490]:
                r9 = (i64 * 9)()
                found = L.ref_frame_seed_rect(int(off), rdlen, 48502, rg, fg, sc.n_ceil(rdlen), 15, r9)
                f2, r = policy.frame_seed_extension_rect(int(off), rdlen, 48502, rg, fg, sc.n_ceil(rdlen))
                assert bool(found) == f2
                assert [r.refl, r.refr, r.refl_pretrim, r.refr_pretrim, r.triml, r.trimr, r.corel, r.corer, r.maxgap] == list(r9)


def test_min_scores_and_intervals():
    sc = policy.Scoring.default(False)
    assert sc.min_score(100) == -60 and sc.min_score(150) == -90 and sc.min_score(50) == -30
    loc = policy.Scoring.default(True)
    assert loc.min_score(300) == 65
    # C2 interval 12, C3 interval 8 (paired), C4 9, C5 18
    assert policy.seed_interval(policy.preset("sensitive").ival, 100) == 12
    assert policy.seed_interval(policy.preset("very-sensitive").ival, 150, True) == 8
    assert policy.seed_interval(policy.preset("very-sensitive", True).ival, 300) == 9
    assert policy.seed_interval(policy.preset("sensitive").ival, 150, True) == 18
    assert policy.n_seeds(100, 22, 12) == 7


def _pe_cases(n=4000, seed=5):
    rng = np.random.default_rng(seed)
    for k in range(n):
        pol = int(rng.integers(1, 5))
        maxfrag = int(rng.choice([50, 200, 500, 800]))
        minfrag = int(rng.choice([0, 0, 30, 150]))
        if minfrag > maxfrag:
            minfrag = 0
        flags = int(rng.integers(0, 32))
        if k % 3 == 0:
            flags = 4 | 8 | 16          # program defaults
        pe = policy.PairedEndPolicy(pol, maxfrag, minfrag, False, bool(flags & 1), bool(flags & 2), bool(flags & 4),
                                    bool(flags & 8), bool(flags & 16))
        len1, len2 = int(rng.integers(20, 260)), int(rng.integers(20, 260))
        reflen = int(rng.choice([300, 2000, 48502]))
        off = int(rng.integers(-40, reflen + 40))
        yield pe, flags, len1, len2, reflen, off, rng


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_other_mate_and_mate_rect_match_reference(lambda_index):
    R = Reference(lambda_index, mirror=False, ref=False)
    L = R.lib
    i64, u64 = C.c_int64, C.c_uint64
    L.ref_frame_mate.argtypes = [C.c_int, u64, u64, C.c_int, C.c_int, C.c_int, i64, i64, u64, u64, u64, u64, u64, i64, u64, C.POINTER(i64)]
    out = (i64 * 15)()
    nfound = 0
    for pe, flags, len1, len2, reflen, off, rng in _pe_cases():
        is1, fw = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        olen = len2 if is1 else len1
        rg, fg = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        maxalcols = olen + rg if rng.integers(0, 4) else -1
        maxns = int(0.15 * olen)
        st = L.ref_frame_mate(pe.pol, pe.maxfrag, pe.minfrag, flags, int(is1), int(fw), off, maxalcols, reflen, len1, len2,
                              rg, fg, maxns, 15, out)
        om = pe.other_mate(is1, fw, off, maxalcols, reflen, len1, len2)
        assert (st != 0) == (om is not None)
        if om is None:
            continue
        oleft, oll, olr, orl, orr, ofw = om
        assert [int(oleft), int(ofw), oll, olr, orl, orr] == list(out[:6])
        found, r = policy.frame_find_mate_rect(not oleft, oll, olr, orl, orr, olen, reflen, rg, fg, maxns, 15)
        assert found == (st == 2)
        assert [r.refl, r.refr, r.refl_pretrim, r.refr_pretrim, r.triml, r.trimr, r.corel, r.corer, r.maxgap] == list(out[6:15])
        nfound += found
    assert nfound > 1000


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_pe_classify_matches_reference(lambda_index):
    R = Reference(lambda_index, mirror=False, ref=False)
    L = R.lib
    i64, u64 = C.c_int64, C.c_uint64
    L.ref_pe_classify.argtypes = [C.c_int, u64, u64, C.c_int, i64, u64, C.c_int, i64, u64, C.c_int]
    seen = set()
    for pe, flags, len1, len2, reflen, off, rng in _pe_cases(6000, seed=8):
        off2 = off + int(rng.integers(-300, 600))
        fw1, fw2 = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        want = L.ref_pe_classify(pe.pol, pe.maxfrag, pe.minfrag, flags, off, len1, int(fw1), off2, len2, int(fw2))
        got = pe.classify_pair(off, len1, fw1, off2, len2, fw2)
        assert got == want
        seen.add(got)
    assert seen == {1, 2, 3, 4, 5}


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("local", [False, True])
def test_mapq_v2_matches_reference(lambda_index, local):
    R = Reference(lambda_index, mirror=False, ref=False)
    L = R.lib
    i64 = C.c_int64
    L.ref_mapq_v2.argtypes = [C.c_void_p, C.c_int, i64, i64, i64, C.c_int, i64]
    sc = policy.Scoring.default(local)
    rng = np.random.default_rng(31)
    seen = set()
    for k in range(6000):
        rdlen = int(rng.choice([50, 100, 150, 250]))
        ordlen = int(rng.choice([0, 0, 100, 150]))
        mn = sc.min_score(rdlen) + (sc.min_score(ordlen) if ordlen else 0)
        pf = sc.perfect_score(rdlen) + (sc.perfect_score(ordlen) if ordlen else 0)
        best = int(rng.integers(mn, pf + 1))
        if k % 5 == 0:
            best = pf
        has_sec = bool(rng.integers(0, 2))
        sec = int(rng.integers(mn, best + 1)) if has_sec else 0
        want = L.ref_mapq_v2(R.h, int(local), rdlen, ordlen, best, int(has_sec), sec)
        got = policy.mapq_v2(best, sec if has_sec else None, mn, pf, monotone=not local)
        assert got == want, (k, rdlen, ordlen, best, has_sec, sec, got, want)
        seen.add(want)
    assert len(seen) > 25


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_random_source_and_rank_seed_hits_match_reference(lambda_index):
    R = Reference(lambda_index, mirror=False, ref=False)
    L = R.lib
    u32p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.ref_rng_draws.argtypes = [C.c_uint32, i32p, u32p, C.c_int, u32p]
    L.ref_rng_draws.restype = None
    L.ref_rank_seed_hits.argtypes = [C.c_uint32, C.c_int, u32p, u32p, C.c_int, u32p, i32p]
    rng = np.random.default_rng(41)
    for trial in range(40):
        seed = int(rng.integers(0, 1 << 32))
        n = 400
        kinds = rng.integers(0, 5, n).astype(np.int32)
        args = rng.integers(1, 1000, n).astype(np.uint32)
        out = np.zeros(n, dtype=np.uint32)
        L.ref_rng_draws(seed, kinds.ctypes.data_as(i32p), args.ctypes.data_as(u32p), n, out.ctypes.data_as(u32p))
        r = policy.RandomSource(seed)
        got = []
        for k, a in zip(kinds, args):
            got.append([r.next_u32, r.next_u2, r.next_bool, lambda: r.next_u32() % int(a), r.next_float_bits][int(k)]())
        assert got == [int(x) for x in out], trial
    for trial in range(300):
        num = int(rng.integers(1, 34))
        fw = (rng.integers(0, 6, num) * (rng.random(num) < 0.6)).astype(np.uint32)
        rc = (rng.integers(0, 400, num) * (rng.random(num) < 0.5)).astype(np.uint32)
        seed = int(rng.integers(0, 1 << 32))
        for all_hits in (False, True):
            oo, of = np.zeros(2 * num, np.uint32), np.zeros(2 * num, np.int32)
            n = L.ref_rank_seed_hits(seed, num, fw.ctypes.data_as(u32p), rc.ctypes.data_as(u32p), int(all_hits),
                                     oo.ctypes.data_as(u32p), of.ctypes.data_as(i32p))
            want = [(int(oo[i]), bool(of[i])) for i in range(n)]
            got = policy.rank_seed_hits([int(x) for x in fw], [int(x) for x in rc], policy.RandomSource(seed), all_hits)
            assert got == want, (trial, all_hits)


def test_gen_rand_seed_known_values():
    # genRandSeed is a static function of pat.cpp; pinned here through values computed by a direct transcription of
    # its three XOR loops (sequence 2 bits at (i & 15) * 2, quality / name bytes at (i & 3) * 8), seed 0
    s = policy.gen_rand_seed([0, 1, 2, 3, 4], [ord(c) for c in "IIIII"], "r1/1", 0)
    base = ((0 + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xffffffff
    x = base
    for i, p in enumerate([0, 1, 2, 3, 4]):
        x ^= p << ((i & 15) << 1)
    for i in range(5):
        x ^= ord("I") << ((i & 3) << 3)
    for i, ch in enumerate(b"r1"):
        x ^= ch << ((i & 3) << 3)
    assert s == x & 0xffffffff


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_device_policy_sources_match_reference_on_host(lambda_index):
    """The __host__ __device__ functions the kernels run (mapq_device.cuh, pe_device.cuh), evaluated on the host through
    bt2g_mapq / bt2g_frame_mate_host / bt2g_pe_classify_host, against the unmodified reference."""
    from bowtie2_b200.lib import MATE_ANCHOR, MATE_FRAME, _PePolicy, _pe_struct, load_library
    lib = load_library()
    R = Reference(lambda_index, mirror=False, ref=False)
    L = R.lib
    i64, u64 = C.c_int64, C.c_uint64
    # MAPQ
    L.ref_mapq_v2.argtypes = [C.c_void_p, C.c_int, i64, i64, i64, C.c_int, i64]
    lib.bt2g_mapq.argtypes = [i64, C.c_int, i64, i64, i64, C.c_int]
    rng = np.random.default_rng(77)
    for local in (False, True):
        sc = policy.Scoring.default(local)
        for k in range(3000):
            rdlen = int(rng.choice([50, 100, 150, 250]))
            ordlen = int(rng.choice([0, 0, 100, 150]))
            mn = sc.min_score(rdlen) + (sc.min_score(ordlen) if ordlen else 0)
            pf = sc.perfect_score(rdlen) + (sc.perfect_score(ordlen) if ordlen else 0)
            best = pf if k % 5 == 0 else int(rng.integers(mn, pf + 1))
            has_sec = bool(rng.integers(0, 2))
            sec = int(rng.integers(mn, best + 1)) if has_sec else 0
            want = L.ref_mapq_v2(R.h, int(local), rdlen, ordlen, best, int(has_sec), sec)
            assert lib.bt2g_mapq(best, int(has_sec), sec, mn, pf, int(not local)) == want
    # mate windows / rectangles and pair classes
    L.ref_frame_mate.argtypes = [C.c_int, u64, u64, C.c_int, C.c_int, C.c_int, i64, i64, u64, u64, u64, u64, u64, i64, u64, C.POINTER(i64)]
    L.ref_pe_classify.argtypes = [C.c_int, u64, u64, C.c_int, i64, u64, C.c_int, i64, u64, C.c_int]
    lib.bt2g_frame_mate_host.argtypes = [C.POINTER(_PePolicy), C.c_void_p, u64, C.c_void_p]
    lib.bt2g_pe_classify_host.argtypes = [C.POINTER(_PePolicy), C.c_void_p, u64, C.c_void_p]
    out15 = (i64 * 15)()
    nfound = 0
    for pe, flags, len1, len2, reflen, off, r2 in _pe_cases(3000, seed=19):
        is1, fw = bool(r2.integers(0, 2)), bool(r2.integers(0, 2))
        olen = len2 if is1 else len1
        rg, fg = int(r2.integers(0, 25)), int(r2.integers(0, 25))
        maxalcols = olen + rg if r2.integers(0, 4) else -1
        a = np.zeros(1, dtype=MATE_ANCHOR)
        a[0] = (off, reflen, len1, len2, maxalcols, rg, fg, int(0.15 * olen), 15, int(is1), int(fw), (0, 0))
        f = np.zeros(1, dtype=MATE_FRAME)
        pp = _pe_struct(pe)
        assert lib.bt2g_frame_mate_host(C.byref(pp), a.ctypes.data_as(C.c_void_p), 1, f.ctypes.data_as(C.c_void_p)) == 0
        st = L.ref_frame_mate(pe.pol, pe.maxfrag, pe.minfrag, flags, int(is1), int(fw), off, maxalcols, reflen, len1, len2, rg, fg,
                              int(0.15 * olen), 15, out15)
        g = f[0]
        assert int(g["status"]) == st
        if st:
            assert [int(g["oleft"]), int(g["ofw"]), int(g["oll"]), int(g["olr"]), int(g["orl"]), int(g["orr"])] == list(out15[:6])
            assert [int(g[k]) for k in ("refl", "refr", "refl_pretrim", "refr_pretrim", "triml", "trimr", "corel", "corer", "maxgap")] == list(out15[6:15])
            nfound += st == 2
        off2 = off + int(r2.integers(-300, 600))
        fw1, fw2 = bool(r2.integers(0, 2)), bool(r2.integers(0, 2))
        pr = np.array([off, len1, int(fw1), off2, len2, int(fw2)], dtype=np.int64)
        cls = np.zeros(1, dtype=np.int32)
        assert lib.bt2g_pe_classify_host(C.byref(pp), pr.ctypes.data_as(C.c_void_p), 1, cls.ctypes.data_as(C.c_void_p)) == 0
        assert int(cls[0]) == L.ref_pe_classify(pe.pol, pe.maxfrag, pe.minfrag, flags, off, len1, int(fw1), off2, len2, int(fw2))
    assert nfound > 800
