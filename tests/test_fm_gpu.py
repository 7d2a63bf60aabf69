"""GPU parity tests for K1/K2: the CUDA path through the C ABI vs the oracle, bit-exact."""
import numpy as np
import pytest

from bowtie2_b200.lib import ReadBatch
from oracle_lib import Oracle

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def seed_interval(ln, const=1.0, coeff=1.15):
    return max(1, int(const + coeff * np.sqrt(ln)))


@pytest.fixture(scope="module", params=["lambda", "synth_small", "synth_large"])
def loaded(request, gpu, lambda_index, synth_index, synth_index_large):
    base = {"lambda": lambda_index, "synth_small": synth_index, "synth_large": synth_index_large}[request.param]
    gpu.load_index_files(base)
    return gpu, Oracle(base), request.param


def test_index_header(loaded):
    gpu, O, _ = loaded
    info, sc = gpu.info(), O.scalars()
    for k in ("len", "bwt_len", "line_rate", "off_rate", "ftab_chars", "num_sides", "side_sz", "side_bwt_sz",
              "n_pat", "n_frag", "offs_len", "ftab_len", "eftab_len", "ebwt_tot_len"):
        assert info[k] == sc[k], k
    assert info["z_off_fw"] == sc["z_off"] and info["z_off_bw"] == O.scalars(True)["z_off"]


def test_rank_lf_ftab(loaded):
    gpu, O, _ = loaded
    n = O.scalars()["bwt_len"]
    rng = np.random.default_rng(5)
    for m in (False, True):
        zo = O.scalars(m)["z_off"]
        side = gpu.info()["side_bwt_len"]
        edge = [0, n - 1, zo, min(zo + 1, n - 1), max(zo, 1) - 1, (zo // side) * side, min(n - 1, (zo // side + 1) * side)]
        rows = np.concatenate([rng.integers(0, n, 4000), edge]).astype(np.uint64)
        assert np.array_equal(gpu.rank4(rows, m), O.rank4(rows, m))
        ch = rng.integers(0, 4, len(rows)).astype(np.uint8)
        assert np.array_equal(gpu.maplf1(rows, ch, m), O.maplf1(rows, ch, m))
        idx = rng.integers(0, O.scalars()["ftab_len"] - 1, 2000).astype(np.uint64)
        assert np.array_equal(gpu.ftab_lohi(idx, m), O.ftab_lohi(idx, m))


def test_maplf_range(loaded):
    """Ebwt::mapLFRange (bt2_idx.h:2268; GroupWalk's step over a range, group_walk.h:897): counts up to the top row, counts
    inside the range and the BWT character of every row, against the oracle (pinned to the reference's own mapLFRange in
    tests/test_oracle.py) and against the identity the reference asserts (:2281-2292) on this library's own rank4."""
    from test_oracle import maplf_range_cases
    gpu, O, _ = loaded
    n = O.scalars()["bwt_len"]
    side = gpu.info()["side_bwt_len"]
    rng = np.random.default_rng(15)
    for m in (False, True):
        zo = O.scalars(m)["z_off"]
        tops, nums = maplf_range_cases(rng, n, zo, side, k=2000)
        upto, inn, chars = gpu.maplf_range(tops, nums, m)
        wu, wi, wc = O.maplf_range(tops, nums, m)
        assert np.array_equal(upto, wu) and np.array_equal(inn, wi) and np.array_equal(chars, wc)
        assert np.array_equal(inn.sum(axis=1), nums)
        inside = tops + nums < n                              # (rank4 takes rows of the BWT)
        bots = gpu.rank4((tops + nums)[inside], m)
        has_z = ((tops <= zo) & (zo < tops + nums))[inside].astype(np.uint64)
        want = bots - upto[inside]
        want[:, 0] += has_z                                   # the "$" row is an A inside a range, and no A for a rank
        assert np.array_equal(inn[inside], want)
    with pytest.raises(RuntimeError):
        gpu.maplf_range([n - 1], [2])
    with pytest.raises(RuntimeError):
        gpu.maplf_range([0], [0])


def test_resolve(loaded):
    gpu, O, _ = loaded
    sc = O.scalars()
    rng = np.random.default_rng(6)
    rows = np.concatenate([rng.integers(0, sc["bwt_len"], 3000), [sc["z_off"], 0, sc["bwt_len"] - 1]]).astype(np.uint64)
    for rej in (False, True):
        hitlen = rng.integers(1, 60, len(rows)).astype(np.uint32)
        joined, tidx, textoff, tlen, flags = gpu.resolve(rows, hitlen, rej)
        want = O.get_offset(rows)
        assert np.array_equal(joined, want)
        for i in range(0, len(rows), 3):
            if int(want[i]) + int(hitlen[i]) > sc["len"]:
                continue
            ok, ti, to, tl, st = O.joined_to_text(int(hitlen[i]), int(want[i]), int(rej))
            assert (flags[i] & 1) == st and ((flags[i] >> 1) & 1) == (0 if ok else 1)
            if ok:
                assert (int(tidx[i]), int(textoff[i]), int(tlen[i])) == (ti, to, tl)


@pytest.mark.parametrize("rate", [0, 1, 2])
def test_resolve_with_dense_sa(loaded, rate):
    """bt2g_build_dense_sa: a denser sample shortens the walk, never changes the offsets."""
    gpu, O, _ = loaded
    sc = O.scalars()
    if rate >= sc["off_rate"]:
        pytest.skip("index already at least this dense")
    rng = np.random.default_rng(16)
    rows = np.concatenate([rng.integers(0, sc["bwt_len"], 3000), [sc["z_off"], 0, sc["bwt_len"] - 1]]).astype(np.uint64)
    hitlen = rng.integers(1, 60, len(rows)).astype(np.uint32)
    base = gpu.resolve(rows, hitlen, False)
    gpu.build_dense_sa(rate)
    try:
        got = gpu.resolve(rows, hitlen, False)
    finally:
        gpu.build_dense_sa(-1)
    assert np.array_equal(got[0], O.get_offset(rows))
    for a, b in zip(base, got):
        assert np.array_equal(a, b)


def _reads_for(name, lambda_reads, synth_genome):
    if name == "lambda":
        return lambda_reads[1][:600]
    from bowtie2_b200 import synth
    reads, _, _ = synth.make_reads(synth_genome, 600, 100, seed=9, sub_rate=0.01, indel_rate=0.001)
    rng = np.random.default_rng(1)
    for r in reads[:60]:
        r[rng.integers(0, len(r))] = 4
    reads.append(np.zeros(1, dtype=np.uint8))            # 1-base read
    reads.append(np.full(30, 4, dtype=np.uint8))          # all-N read
    reads.append(reads[0][:9].copy())                     # shorter than ftabChars
    return reads


def test_exact_sweep(loaded, lambda_reads, synth_genome):
    gpu, O, name = loaded
    reads = _reads_for(name, lambda_reads, synth_genome)
    mine, ee = gpu.exact_sweep(ReadBatch.from_list(reads))
    for i, r in enumerate(reads):
        nelt, m, tb = O.exact_sweep(r)
        assert list(mine[i]) == m and [int(x) for x in ee[i]] == tb, i
    # strand switches
    mine, ee = gpu.exact_sweep(ReadBatch.from_list(reads[:50]), nofw=True)
    for i, r in enumerate(reads[:50]):
        nelt, m, tb = O.exact_sweep(r, nofw=True)
        assert list(mine[i]) == m and [int(x) for x in ee[i]] == tb


@pytest.mark.parametrize("L,off", [(22, 0), (20, 3), (10, 1), (32, 0)])
def test_seed_search(loaded, lambda_reads, synth_genome, L, off):
    gpu, O, name = loaded
    reads = [r for r in _reads_for(name, lambda_reads, synth_genome) if len(r) >= L + off]
    batch = ReadBatch.from_list(reads)
    interval = np.array([seed_interval(len(r)) for r in reads], dtype=np.int32)
    MS = 32
    out, ns = gpu.seed_search(batch, L, interval, off, MS)
    nhit = 0
    for i, r in enumerate(reads):
        n, want = O.seed_search(r, L, int(interval[i]), off, MS)
        assert n == ns[i]
        assert np.array_equal(out[i], want), (i, L, off)
        nhit += int((want[:, :, 1] > want[:, :, 0]).sum())
    assert nhit > 0


@pytest.mark.parametrize("k", [11, 12, 13])
def test_seed_search_with_extended_table(loaded, lambda_reads, synth_genome, k):
    """bt2g_build_seed_table: the k-mer start table changes where the search starts, never what it returns."""
    gpu, O, name = loaded
    reads = _reads_for(name, lambda_reads, synth_genome)
    gpu.build_seed_table(k)
    try:
        for L, off in ((22, 0), (20, 3), (13, 1), (12, 0), (32, 2)):
            rr = [r for r in reads if len(r) >= L + off]
            batch = ReadBatch.from_list(rr)
            interval = np.array([seed_interval(len(r)) for r in rr], dtype=np.int32)
            out, ns = gpu.seed_search(batch, L, interval, off, 32)
            nhit = 0
            for i, r in enumerate(rr):
                n, want = O.seed_search(r, L, int(interval[i]), off, 32)
                assert n == ns[i]
                assert np.array_equal(out[i], want), (i, L, off, k)
                nhit += int((want[:, :, 1] > want[:, :, 0]).sum())
            assert nhit > 0
    finally:
        gpu.build_seed_table(0)


def test_get_stretch(loaded, synth_genome):
    gpu, O, name = loaded
    nref = 1 if name == "lambda" else len(synth_genome)
    rng = np.random.default_rng(2)
    tidx = rng.integers(0, nref, 200)
    off = rng.integers(-40, 48000 if name == "lambda" else 40010, 200)
    cnt = rng.integers(1, 300, 200)
    out = gpu.get_stretch(tidx, off, cnt, 300)
    for i in range(200):
        assert np.array_equal(out[i, :cnt[i]], O.get_stretch(int(tidx[i]), int(off[i]), int(cnt[i]))), i
