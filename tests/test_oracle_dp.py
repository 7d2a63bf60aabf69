"""CPU: the plain-C DP restatement (with the reference's full mask / branch-stack backtrace) vs the
unmodified reference SwAligner."""
import numpy as np
import pytest

from bowtie2_b200 import policy, synth
from oracle_lib import Oracle, Reference, have_reference, oracle_dp, ref_dp


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("rdlen,sub,indel", [(100, 0.015, 0.003), (60, 0.03, 0.01), (180, 0.01, 0.004)])
def test_oracle_dp_matches_reference(synth_index, synth_genome, rdlen, sub, indel, local):
    O, R = Oracle(synth_index), Reference(synth_index)
    sc = policy.Scoring.default(local)
    reads, quals, truth = synth.make_reads(synth_genome, 120, rdlen, seed=3 * rdlen, sub_rate=sub, indel_rate=indel, random_frac=0.05)
    rng = np.random.default_rng(rdlen)
    for r in reads[:15]:
        r[rng.integers(0, len(r))] = 4
    if local:
        for r in reads[::3]:
            k = int(rng.integers(3, 12))
            r[:k] = rng.integers(0, 4, k)
    n = nfound = 0
    for i, (r, q, (c, p, strand)) in enumerate(zip(reads, quals, truth)):
        if c < 0:
            c, p, strand = 0, int(rng.integers(0, 30000)), 1
        for bump in (0, 25):
            minsc = sc.min_score(rdlen) + bump
            off = p + int(rng.integers(-3, 4))
            tlen = len(synth_genome[c])
            found, rect = policy.frame_seed_extension_rect(off, rdlen, tlen, sc.max_read_gaps(minsc, rdlen),
                                                           sc.max_ref_gaps(minsc, rdlen), sc.n_ceil(rdlen))
            if not found:
                continue
            if minsc > sc.perfect_score(rdlen):
                continue
            want = ref_dp(R, local, r, q, strand > 0, c, tlen, rect, minsc, max_cands=8192)
            got = oracle_dp(O, local, r, q, strand > 0, c, rect, minsc, sc.n_ceil_raw(rdlen), max_cands=8192)
            n += 1
            assert got["found"] == want["found"], (i, got, want["found"])
            if not want["found"]:
                assert got["best"] < minsc
                continue
            nfound += 1
            assert got["best"] == want["best"] and got["ncand"] == want["ncand"] and got["cands"] == want["cands"]
            assert got["naln"] == want["naln"]
            for a, b in zip(got["alns"], want["alns"]):
                assert a == b, (i, a, b)
    assert nfound > 100
