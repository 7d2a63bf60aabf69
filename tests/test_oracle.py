"""CPU tests: the oracle restatement against the reference's own answers (golden fixtures and,
where oracle/_ref is built, the live reference), and the C-ABI library's symbol table."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle_lib import Oracle, Reference, have_reference, ROOT


def seed_interval(ln, const=1.0, coeff=1.15):
    return max(1, int(const + coeff * np.sqrt(ln)))


def maplf_range_cases(rng, n, zo, side_len, k=300):
    """[top, top+num) ranges for Ebwt::mapLFRange: random narrow ranges (what GroupWalk holds), ranges over one and several
    side ends, one row, the rows around the "$" row (the reference tallies it as an A inside a range), the last rows."""
    tops = [int(t) for t in rng.integers(0, n - 1, k)]
    nums = [int(x) for x in rng.integers(1, 40, k)]
    edge = [(0, 1), (0, side_len), (side_len - 1, 2), (side_len - 3, 2 * side_len + 7), (n - 1, 1), (max(0, n - 50), min(50, n)),
            (zo, 1), (max(zo, 5) - 5, 11), (min(zo + 1, n - 1), 1), (side_len * (zo // side_len), side_len), (3, min(n - 3, 5 * side_len + 1))]
    for t, m in edge:
        tops.append(t); nums.append(m)
    nums = [max(1, min(m, n - t)) for t, m in zip(tops, nums)]
    return np.array(tops, dtype=np.uint64), np.array(nums, dtype=np.uint64)


def test_oracle_matches_golden_fm(lambda_index, golden_fm):
    g = golden_fm
    O = Oracle(lambda_index)
    for tag, m in (("fw", False), ("bw", True)):
        rows = g[f"rank_rows_{tag}"]
        assert np.array_equal(O.rank4(rows, m), g[f"rank4_{tag}"])
        assert np.array_equal(O.maplf1(rows, g[f"lf1_chars_{tag}"], m), g[f"lf1_{tag}"])
        assert np.array_equal(O.ftab_lohi(g[f"ftab_idx_{tag}"], m), g[f"ftab_{tag}"])
    assert np.array_equal(O.get_offset(g["off_rows"]), g["offsets"])
    for off, qlen, ok, ti, to, tl, st in g["joined"]:
        got = O.joined_to_text(int(qlen), int(off), 1)
        assert tuple(int(x) for x in got) == (int(ok), int(ti), int(to), int(tl), int(st))


def test_oracle_matches_golden_seeds(lambda_index, golden_fm, lambda_reads):
    g = golden_fm
    O = Oracle(lambda_index)
    names, reads, quals = lambda_reads
    L, MS = (int(x) for x in g["seed_params"])
    for i in range(len(g["sweep"])):
        codes = reads[i]
        nelt, mine, tb = O.exact_sweep(codes)
        assert [nelt] + mine + tb == [int(x) for x in g["sweep"][i]], i
        n, out = O.seed_search(codes, L, seed_interval(len(codes)), 0, MS)
        assert np.array_equal(out, g["seeds"][i]), i


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("which", ["small", "large"])
def test_oracle_matches_live_reference(which, synth_index, synth_index_large, synth_genome):
    base = synth_index if which == "small" else synth_index_large
    O = Oracle(base)
    R = Reference(base, large=(which == "large"))
    so, sr = O.scalars(), R.scalars()
    assert so == sr
    assert O.scalars(True) == R.scalars(True)
    n = so["bwt_len"]
    rng = np.random.default_rng(3)
    for m in (False, True):
        zo = O.scalars(m)["z_off"]
        rows = np.concatenate([rng.integers(0, n, 1500), [0, n - 1, zo, min(zo + 1, n - 1)]]).astype(np.uint64)
        assert np.array_equal(O.rank4(rows, m), R.rank4(rows, m))
        ch = rng.integers(0, 4, len(rows))
        assert np.array_equal(O.maplf1(rows, ch, m), R.maplf1(rows, ch, m))
        idx = rng.integers(0, so["ftab_len"] - 1, 1000)
        assert np.array_equal(O.ftab_lohi(idx, m), R.ftab_lohi(idx, m))
        tops, nums = maplf_range_cases(rng, n, zo, so["side_bwt_sz"] * 4)
        for got, want in zip(O.maplf_range(tops, nums, m), R.maplf_range(tops, nums, m)):
            assert np.array_equal(got, want)
    rows = rng.integers(0, n, 1000)
    offs = O.get_offset(rows)
    assert np.array_equal(offs, R.get_offset(rows))
    for off in offs[:400]:
        for qlen in (1, 30):
            for rej in (0, 1):
                if int(off) + qlen <= so["len"]:
                    assert O.joined_to_text(qlen, int(off), rej) == R.joined_to_text(qlen, int(off), rej)
    # reference windows including N gaps and off-end padding
    for t in range(len(synth_genome)):
        for off in (-5, 0, 19990, 39990):
            assert np.array_equal(O.get_stretch(t, off, 64), R.get_stretch(t, off, 64))
    # reads with mismatches / Ns: exact sweep + multiseed
    from bowtie2_b200 import synth
    reads, quals, _ = synth.make_reads(synth_genome, 300, 75, seed=5, sub_rate=0.02, indel_rate=0.002)
    for r in reads[:50]:
        r[rng.integers(0, len(r))] = 4
    for r in reads:
        assert O.exact_sweep(r) == R.exact_sweep(r)
        for L, off in ((20, 0), (22, 3), (10, 1), (32, 0)):
            no, oo = O.seed_search(r, L, seed_interval(len(r)), off, 32)
            nr, orr = R.seed_search(r, L, seed_interval(len(r)), off, 32)
            assert no == nr and np.array_equal(oo, orr)


def test_cabi_library_exports_every_declared_symbol():
    """The C-ABI shared library loads and exports every symbol include/bt2g.h declares."""
    from bowtie2_b200.lib import EXPORTS, library_path
    hdr = open(os.path.join(ROOT, "include", "bt2g.h")).read()
    declared = set(re.findall(r"\b(bt2g_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    path = library_path()
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", path]).decode()
    exported = set(re.findall(r"\bT (bt2g_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    # no oracle symbols may be linked into the product
    assert "bt2o_" not in out and "ref_open" not in out


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bowtie2_b200 import Bt2Gpu, Bt2GpuError
    with pytest.raises(Bt2GpuError):
        Bt2Gpu(0)
