"""csrc/policy_engine.cpp (bt2g_policy_align): the exact search policy as C++20 coroutines scheduled in waves over the
entry-point table.  With the table answered by the CPU oracle behind the entry points' own array conventions
(tests/fake_gpu.py) the SAM must be identical to the reference program's golden files, and the result arrays identical to
the Python engine's (the pinned specification) across modes and options."""
import os
import subprocess

import numpy as np
import pytest

from bowtie2_b200 import policy, synth
from bowtie2_b200.lib import PAIR_RESULT, READ_RESULT, ReadBatch, load_library, policy_align, policy_params, sam_format
from bowtie2_b200.policy_engine import PairedPolicyEngine, PolicyEngine
from conftest import GOLDEN, read_fastq_codes
from fake_gpu import FakeGpu, backend_table
from oracle_lib import Oracle, have_reference, ref_bin
from policy_backend import OracleBackend
from test_policy_engine import _fill


def _table(index, local=False, off_size=4):
    fake = FakeGpu(Oracle(index), off_size)
    fake.set_scoring(local)
    be, keep = backend_table(fake)
    return be, keep, fake


@pytest.mark.parametrize("fixture,index,ref_names,local", [
    ("lambda_U_sensitive", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], False),
    ("lambda_U_local", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], True),
    ("rep_U_sensitive", "rep_index", ["ctg1", "ctg2"], False),
])
def test_unpaired_sam_identical_to_golden(fixture, index, ref_names, local, request):
    base = request.getfixturevalue(index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, fixture + ".sam")) if not l.startswith("@")]
    names, reads, quals = read_fastq_codes(os.path.join(GOLDEN, fixture.split("_")[0] + "_reads_1.fq"), len(golden))
    be, keep, fake = _table(base, local)
    lib = load_library()
    batch = ReadBatch.from_list(reads, quals)
    res, ops, _, (waves, calls, requests) = policy_align(lib, be, policy_params("sensitive", local=local), batch, names)
    lines = sam_format(lib, batch, res, ops, ref_names, read_names=names, local=local).rstrip("\n").split("\n")
    assert lines == golden
    assert waves < 80 and calls * 20 < requests          # batched: one entry-point call per primitive and wave


@pytest.mark.parametrize("fixture,index,ref_names", [
    ("lambda", "lambda_index", ["gi|9626243|ref|NC_001416.1|"]),
    ("rep", "rep_index", ["ctg1", "ctg2"]),
])
def test_paired_sam_identical_to_golden(fixture, index, ref_names, request):
    base = request.getfixturevalue(index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, f"{fixture}_P_sensitive.sam")) if not l.startswith("@")]
    n = len(golden) // 2
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_1.fq"), n)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_2.fq"), n)
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    R, Q, N = il(r1, r2), il(q1, q2), il(n1, n2)
    be, keep, fake = _table(base)
    lib = load_library()
    batch = ReadBatch.from_list(R, Q)
    res, ops, pairs, stats = policy_align(lib, be, policy_params("sensitive", paired=True, max_inflight=64), batch, N)
    lines = sam_format(lib, batch, res, ops, ref_names, read_names=N, pairs=pairs).rstrip("\n").split("\n")
    assert lines == golden


def _python_results(index, reads, quals, names, paired, local, off_size, scoring=None, **kw):
    n = len(reads)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, max(len(r) for r in reads) + 64), dtype=np.uint8)
    backend = OracleBackend(Oracle(index), off_size=off_size, local=local, scoring=scoring)
    if not paired:
        eng = PolicyEngine(backend, "sensitive", local=local, **kw)
        for i in range(n):
            r = eng.align_read(reads[i], quals[i], names[i])
            if r.aligned:
                _fill(res, ops, i, r, reads[i])
        return res, ops, None
    eng = PairedPolicyEngine(backend, "sensitive", local=local, **kw)
    pairs = np.zeros(n // 2, dtype=PAIR_RESULT)
    for i in range(n // 2):
        pr = eng.align_pair(reads[2 * i], quals[2 * i], names[2 * i], reads[2 * i + 1], quals[2 * i + 1], names[2 * i + 1])
        pairs[i]["pair_type"] = pr.pair_type
        for k in range(2):
            if pr.mates[k].aligned:
                _fill(res, ops, 2 * i + k, pr.mates[k], reads[2 * i + k])
    return res, ops, pairs


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("paired,local,large,kw", [
    (False, False, False, {}),
    (False, True, False, {}),
    (True, False, False, {}),
    (True, True, False, {}),
    (True, False, True, {}),                                   # .bt2l: 64-bit RNG draws
    (False, False, False, dict(k=4)),                          # -k: primary records (the secondaries are the Python engine's)
    (True, False, False, dict(all_hits=True)),
    (True, False, False, dict(mixed=False, discord=False)),
    (False, False, False, dict(nofw=True, seed=9)),
])
def test_same_results_as_the_python_engine(tmp_path, paired, local, large, kw):
    """repeat-rich synthetic genome; results (locus, strand, scores, MAPQ, trims, op strings, pair types) equal to the pinned
    Python engine's, through the same oracle"""
    genome = synth.make_genome(n_contigs=3, contig_len=60000, seed=11, repeat_frac=0.5, repeat_len=250, repeat_copies=150, n_gap=37)
    fa, base = str(tmp_path / "g.fa"), str(tmp_path / "g")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-" + ("l" if large else "s")), "--seed", "0", "--quiet", fa, base])
    if paired:
        reads, quals, _ = synth.make_pairs(genome, 120, 100, seed=41, sub_rate=0.02, indel_rate=0.003, hard_frac=0.2, hard_period=12, ins_sd=100)
        names = [f"r{i // 2}" for i in range(len(reads))]
    else:
        reads, quals, _ = synth.make_reads(genome, 250, 100, seed=42, sub_rate=0.02, indel_rate=0.003)
        names = [f"r{i}" for i in range(len(reads))]
    off_size = 8 if large else 4
    want_res, want_ops, want_pairs = _python_results(base, reads, quals, names, paired, local, off_size, **kw)
    be, keep, fake = _table(base, local, off_size)
    pk = dict(kw)
    prm = policy_params("sensitive", local=local, paired=paired, k=pk.pop("k", None), all_hits=pk.pop("all_hits", False),
                        nofw=pk.pop("nofw", False), seed=pk.pop("seed", 0), mixed=pk.pop("mixed", True), discord=pk.pop("discord", True))
    assert not pk
    res, ops, pairs, stats = policy_align(load_library(), be, prm, ReadBatch.from_list(reads, quals), names)
    for f in ("found", "score", "score2", "fw", "tidx", "refoff", "nops", "trim_left", "trim_right", "mapq", "pad"):
        assert np.array_equal(res[f], want_res[f]), (f, np.nonzero(res[f] != want_res[f])[0][:5])
    w = want_ops.shape[1]
    assert np.array_equal(ops[:, :w], want_ops)
    if paired:
        assert np.array_equal(pairs["pair_type"], want_pairs["pair_type"])
    assert (res["found"] != 0).mean() > 0.3


@pytest.mark.parametrize("fixture,index,ref_names,local,paired", [
    ("lambda_U_sensitive", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], False, False),
    ("lambda_U_local", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], True, False),
    ("rep_P_sensitive", "rep_index", ["ctg1", "ctg2"], False, True),
])
def test_whole_path_in_c_over_the_oracle_table(fixture, index, ref_names, local, paired, request):
    """the compiled engine over oracle/bt2_oracle_table.c (the entry points answered by the plain-C restatement): the whole path
    in C / C++ on the CPU, byte-identical to the reference program; also pins the stand-in device of tests/fake_gpu.py, which
    implements the same conventions in Python"""
    from oracle_lib import oracle_policy_table
    base = request.getfixturevalue(index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, fixture + ".sam")) if not l.startswith("@")]
    pre = fixture.split("_")[0]
    n = len(golden) // (2 if paired else 1)
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, pre + "_reads_1.fq"), n)
    if paired:
        n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, pre + "_reads_2.fq"), n)
        il = lambda a, b: [x for p in zip(a, b) for x in p]
        R, Q, N = il(r1, r2), il(q1, q2), il(n1, n2)
    else:
        R, Q, N = r1, q1, n1
    be, keep = oracle_policy_table(Oracle(base), local)
    lib = load_library()
    batch = ReadBatch.from_list(R, Q)
    res, ops, pairs, stats = policy_align(lib, be, policy_params("sensitive", local=local, paired=paired, host_threads=3), batch, N)
    lines = sam_format(lib, batch, res, ops, ref_names, read_names=N, pairs=pairs, local=local).rstrip("\n").split("\n")
    assert lines == golden
    # the Python stand-in device gives the same arrays
    be2, keep2, fake = _table(base, local)
    res2, ops2, pairs2, _ = policy_align(lib, be2, policy_params("sensitive", local=local, paired=paired), batch, N)
    assert res.tobytes() == res2.tobytes() and ops.tobytes() == ops2.tobytes()


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,kw,cap", [(["-k", "3"], dict(k=3), 3), (["-k", "12"], dict(k=12), 12), (["-a"], dict(all_hits=True), 400)])
def test_k_and_all_modes_unpaired_compiled(tmp_path, args, kw, cap):
    """bt2g_policy_align_k: every reported alignment of -k N / -a (primary, then the secondaries in the reference's order):
    SAM identical to the reference program's"""
    from bowtie2_b200.lib import policy_align_k
    genome = synth.make_genome(n_contigs=3, contig_len=60000, seed=11, repeat_frac=0.5, repeat_len=250, repeat_copies=150, n_gap=37)
    fa, base = str(tmp_path / "g.fa"), str(tmp_path / "g")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", fa, base])
    reads, quals, _ = synth.make_reads(genome, 250, 100, seed=79, sub_rate=0.02, indel_rate=0.003)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, reads, quals)
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-U", fq] + args,
                                  stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    names = [f"r{i}" for i in range(len(reads))]
    be, keep, fake = _table(base)
    prm = policy_params("sensitive", k=kw.get("k"), all_hits=kw.get("all_hits", False))
    res, ops, cnt, truncated, stats = policy_align_k(load_library(), be, prm, ReadBatch.from_list(reads, quals), names, cap)
    assert not truncated
    rows = [(i, j) for i in range(len(reads)) for j in range(max(int(cnt[i]), 1))]        # an unaligned read still prints one record
    R, Q, N = [reads[i] for i, _ in rows], [quals[i] for i, _ in rows], [names[i] for i, _ in rows]
    res_f = np.array([res[i, j] for i, j in rows], dtype=READ_RESULT)
    ops_f = np.stack([ops[i, j] for i, j in rows])
    ref_names = [l.split("\t")[1][3:] for l in out.split("\n") if l.startswith("@SQ")]
    lines = sam_format(load_library(), ReadBatch.from_list(R, Q), res_f, ops_f, ref_names, read_names=N).rstrip("\n").split("\n")
    assert lines == want, next((a, b) for a, b in zip(lines, want) if a != b)
    assert sum(int(l.split("\t")[1]) & 256 != 0 for l in want) > 20
    # a cap below the number of alignments drops the extra ones and says so
    res2, ops2, cnt2, truncated2, _ = policy_align_k(load_library(), be, prm, ReadBatch.from_list(reads, quals), names, 2)
    assert truncated2 and int(cnt2.max()) == 2 and np.array_equal(res2[:, 0], res[:, 0])


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,kw,cap", [(["-k", "3"], dict(k=3), 3), (["-a"], dict(all_hits=True), 64)])
def test_paired_k_and_all_modes_equal_the_reference_program(tmp_path, args, kw, cap):
    """bt2g_policy_align_pairs_k: every record of paired -k N / -a -- the primaries, the further concordant pairs, the mates' further
    unpaired alignments beside the opposite mate's primary -- formatted by bt2g_sam_format straight from the entry rows (bit 9 of
    `found` = mate context only, not printed): identical to the reference program's SAM"""
    from bowtie2_b200.lib import policy_align_pairs_k
    from test_policy_engine import _synth_index
    genome, base = _synth_index(tmp_path)
    n = 150
    reads, quals, _ = synth.make_pairs(genome, n, 100, seed=32, sub_rate=0.02, indel_rate=0.003, hard_frac=0.2, hard_period=12,
                                       ins_mean=300, ins_sd=90)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, reads[0::2], quals[0::2])
    synth.write_fastq(f2, reads[1::2], quals[1::2])
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-1", f1, "-2", f2] + args,
                                  stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    names = [f"r{i // 2}" for i in range(2 * n)]
    be, keep, fake = _table(base)
    prm = policy_params("sensitive", paired=True, k=kw.get("k"), all_hits=kw.get("all_hits", False))
    res, ops, pairs, cnt, truncated, stats = policy_align_pairs_k(load_library(), be, prm, ReadBatch.from_list(reads, quals), names, cap)
    assert not truncated and int(cnt.max()) > 1
    # flatten the used entries: the pair's reads repeated per entry
    R, Q, N, rr, oo, pp = [], [], [], [], [], []
    for i in range(n):
        for e in range(int(cnt[i])):
            R += [reads[2 * i], reads[2 * i + 1]]; Q += [quals[2 * i], quals[2 * i + 1]]; N += [names[2 * i], names[2 * i + 1]]
            rr.append(res[i, e]); oo.append(ops[i, e]); pp.append(pairs[i, e])
    lines = sam_format(load_library(), ReadBatch.from_list(R, Q), np.concatenate(rr), np.concatenate(oo), ["chr1", "chr2", "chr3"], read_names=N,
                       pairs=np.array(pp, dtype=PAIR_RESULT)).rstrip("\n").split("\n")
    diff = [(a, b) for a, b in zip(lines, want) if a != b]
    assert len(lines) == len(want) and not diff, (len(lines), len(want), diff[:1])
    assert sum(int(l.split("\t")[1]) & 256 != 0 for l in want) > 20
