"""GPU: paired-end framing kernels (bt2g_frame_mate, bt2g_pe_classify) against the host mirror
bowtie2_b200.policy, which tests/test_policy.py pins against the unmodified reference
(PairedEndPolicy::otherMate, peClassifyPair, DynProgFramer::frameFindMateRect)."""
import numpy as np
import pytest

from bowtie2_b200 import policy
from bowtie2_b200.lib import MATE_ANCHOR

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _policies():
    yield policy.PairedEndPolicy()                       # program defaults
    rng = np.random.default_rng(3)
    for _ in range(24):
        fl = int(rng.integers(0, 32))
        maxfrag = int(rng.choice([50, 200, 500, 800]))
        minfrag = int(rng.choice([0, 30, 150]))
        if minfrag > maxfrag:
            minfrag = 0
        yield policy.PairedEndPolicy(int(rng.integers(1, 5)), maxfrag, minfrag, False, bool(fl & 1), bool(fl & 2),
                                     bool(fl & 4), bool(fl & 8), bool(fl & 16))


def test_frame_mate_gpu_vs_policy(gpu):
    rng = np.random.default_rng(17)
    nfound = 0
    for pe in _policies():
        n = 500
        a = np.zeros(n, dtype=MATE_ANCHOR)
        a["reflen"] = rng.choice([300, 2000, 48502], size=n)
        a["off"] = rng.integers(-40, 2200, size=n)
        a["len1"] = rng.integers(20, 260, size=n)
        a["len2"] = rng.integers(20, 260, size=n)
        a["is1"] = rng.integers(0, 2, size=n)
        a["fw"] = rng.integers(0, 2, size=n)
        a["maxrdgap"] = rng.integers(0, 25, size=n)
        a["maxrfgap"] = rng.integers(0, 25, size=n)
        olen = np.where(a["is1"] != 0, a["len2"], a["len1"]).astype(np.int64)
        a["maxalcols"] = np.where(rng.integers(0, 4, size=n) != 0, olen + a["maxrdgap"], -1)
        a["maxns"] = (0.15 * olen).astype(np.int32)
        a["maxhalf"] = 15
        got = gpu.frame_mate(pe, a)
        for i in range(n):
            x = a[i]
            om = pe.other_mate(bool(x["is1"]), bool(x["fw"]), int(x["off"]), int(x["maxalcols"]), int(x["reflen"]),
                               int(x["len1"]), int(x["len2"]))
            g = got[i]
            if om is None:
                assert g["status"] == 0
                continue
            oleft, oll, olr, orl, orr, ofw = om
            assert (bool(g["oleft"]), bool(g["ofw"]), int(g["oll"]), int(g["olr"]), int(g["orl"]), int(g["orr"])) == \
                   (oleft, ofw, oll, olr, orl, orr)
            found, r = policy.frame_find_mate_rect(not oleft, oll, olr, orl, orr, int(olen[i]), int(x["reflen"]),
                                                   int(x["maxrdgap"]), int(x["maxrfgap"]), int(x["maxns"]), 15)
            assert g["status"] == (2 if found else 1)
            assert [int(g[k]) for k in ("refl", "refr", "refl_pretrim", "refr_pretrim", "triml", "trimr", "corel", "corer", "maxgap")] == \
                   [r.refl, r.refr, r.refl_pretrim, r.refr_pretrim, r.triml, r.trimr, r.corel, r.corer, r.maxgap]
            nfound += found
    assert nfound > 5000


def test_pe_classify_gpu_vs_policy(gpu):
    rng = np.random.default_rng(23)
    seen = set()
    for pe in _policies():
        n = 800
        p = np.zeros((n, 6), dtype=np.int64)
        p[:, 0] = rng.integers(0, 3000, size=n)
        p[:, 1] = rng.integers(20, 260, size=n)
        p[:, 2] = rng.integers(0, 2, size=n)
        p[:, 3] = p[:, 0] + rng.integers(-300, 600, size=n)
        p[:, 4] = rng.integers(20, 260, size=n)
        p[:, 5] = rng.integers(0, 2, size=n)
        got = gpu.pe_classify(pe, p)
        for i in range(n):
            want = pe.classify_pair(int(p[i, 0]), int(p[i, 1]), bool(p[i, 2]), int(p[i, 3]), int(p[i, 4]), bool(p[i, 5]))
            assert int(got[i]) == want
            seen.add(want)
    assert seen == {1, 2, 3, 4, 5}
