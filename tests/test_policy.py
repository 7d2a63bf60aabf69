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
