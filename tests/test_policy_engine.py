"""bowtie2_b200.policy_engine: the reference's sequential, RNG-driven search policy for unpaired end-to-end reads,
replayed over the CPU oracle's primitives, must reproduce the reference PROGRAM byte for byte -- every SAM field
including MAPQ and XS:i, the locus chosen among equal repeats -- and its per-read work counters (--read-times:
ZI extend-loop iterations, XD gapped DPs, XU ungapped extensions, YR redundant seed hits), which pins the order and
number of RNG draws."""
import os
import subprocess

import numpy as np
import pytest

from bowtie2_b200 import synth
from bowtie2_b200.lib import READ_RESULT, ReadBatch, load_library, sam_format
from bowtie2_b200.policy_engine import PolicyEngine, Random1toN, aln_to_ops
from bowtie2_b200.policy import RandomSource
from conftest import GOLDEN, read_fastq_codes
from oracle_lib import Oracle, have_reference, ref_bin
from policy_backend import OracleBackend

KEEP = ("AS", "XS", "XN", "XM", "XO", "XG", "NM", "MD", "YS", "YT", "YF")


def _fill(res, ops, j, r, read):
    a = r.aln
    o = aln_to_ops(a, read)
    res[j]["found"] = 2 if (not a.edits and a.ext == a.rdlen) else 1
    res[j]["score"] = a.score
    if r.xs is not None:
        res[j]["score2"] = r.xs
    res[j]["fw"] = int(a.fw); res[j]["tidx"] = a.tidx; res[j]["refoff"] = a.refoff; res[j]["nops"] = len(o)
    res[j]["trim_left"] = a.trim_left; res[j]["trim_right"] = a.rdlen - a.ext - a.trim_left
    res[j]["mapq"] = r.mapq; res[j]["pad"] = a.refns
    ops[j, :len(o)] = o


def _run_engine(index, reads, quals, names, ref_names, preset, local=False):
    O = Oracle(index)
    eng = PolicyEngine(OracleBackend(O, local=local), preset, local=local)
    n = len(reads)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, max(len(r) for r in reads) + 64), dtype=np.uint8)
    outs = []
    for i in range(n):
        r = eng.align_read(reads[i], quals[i], names[i])
        outs.append(r)
        if r.aligned:
            _fill(res, ops, i, r, reads[i])
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ref_names, read_names=names).rstrip("\n").split("\n")
    return lines, outs


@pytest.mark.parametrize("fixture,index,ref_names", [
    ("lambda", "lambda_index", ["gi|9626243|ref|NC_001416.1|"]),
    ("rep", "rep_index", ["ctg1", "ctg2"]),
    ("lambda:local", "lambda_index", ["gi|9626243|ref|NC_001416.1|"]),
])
def test_sam_identical_to_golden(fixture, index, ref_names, request):
    base = request.getfixturevalue(index)
    local = fixture.endswith(":local")
    fixture = fixture.split(":")[0]
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, f"{fixture}_U_{'local' if local else 'sensitive'}.sam")) if not l.startswith("@")]
    names, reads, quals = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_1.fq"), len(golden))
    lines, outs = _run_engine(base, reads, quals, names, ref_names, "sensitive", local)
    bad = [i for i in range(len(golden)) if lines[i] != golden[i]]
    assert not bad, (len(bad), lines[bad[0]], golden[bad[0]])
    if fixture == "rep":
        assert sum(o.n_alns >= 2 for o in outs) > 80          # ties in repeats were decided by the replayed RNG


def _reference_run(index, fq, preset, local=False):
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), *(["--local"] if local else []), "--" + preset, "--seed", "0", "-p", "1", "--read-times", "-x", index, "-U", fq],
                                  stderr=subprocess.DEVNULL).decode()
    full = [l for l in out.split("\n") if l and not l.startswith("@")]
    recs, counters = [], []
    for l in full:
        f = l.split("\t")
        recs.append("\t".join(f[:11] + [x for x in f[11:] if x[:2] in KEEP]))
        tags = {t[:2]: t[5:] for t in f[11:]}
        counters.append(tuple(int(tags[k]) for k in ("ZI", "XD", "XU", "YR")) if "ZI" in tags else None)
    return recs, counters


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("preset,rdlen,n,genome_kw", [
    ("local:sensitive", 100, 300, {}),
    ("local:very-sensitive", 300, 100, {}),               # configuration 4's preset and read length (--local)
    ("local:sensitive", 100, 200, dict(contig_len=120000, repeat_frac=0.6, repeat_len=250, repeat_copies=400)),
    ("sensitive", 100, 500, {}),
    ("very-sensitive", 150, 300, {}),
    ("fast", 50, 400, {}),
    ("very-fast", 250, 200, {}),
    ("sensitive", 36, 400, {}),
    # half of the genome in 400-copy repeat families: -M ceiling, weighted range sampling, re-seeding rounds
    ("sensitive", 100, 400, dict(contig_len=120000, repeat_frac=0.6, repeat_len=250, repeat_copies=400)),
])
def test_sam_and_work_counters_identical_to_reference_program(tmp_path, preset, rdlen, n, genome_kw):
    local = preset.startswith("local:")
    preset = preset.split(":")[-1]
    kw = dict(n_contigs=3, contig_len=40000, seed=11, repeat_frac=0.05, repeat_len=300, repeat_copies=12, n_gap=37)
    kw.update(genome_kw)
    genome = synth.make_genome(**kw)
    fa, base, fq = str(tmp_path / "g.fa"), str(tmp_path / "g"), str(tmp_path / "r.fq")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", fa, base])
    reads, quals, _ = synth.make_reads(genome, n, rdlen, seed=5 + rdlen, sub_rate=0.02, indel_rate=0.003)
    synth.write_fastq(fq, reads, quals)
    want, counters = _reference_run(base, fq, preset, local)
    names = [f"r{i}" for i in range(len(reads))]
    lines, outs = _run_engine(base, reads, quals, names, [f"chr{k + 1}" for k in range(len(genome))], preset, local)
    bad = [i for i in range(len(want)) if lines[i] != want[i]]
    assert not bad, (len(bad), lines[bad[0]], want[bad[0]])
    for i, o in enumerate(outs):
        if o.counters is not None and counters[i] is not None:
            c = o.counters
            assert (c["ZI"], c["XD"], c["XU"], c["YR"]) == counters[i], (i, c, counters[i])
    if local:
        assert sum("S" in l.split("\t")[5] for l in want) > 10          # soft-clipped records were compared
    if genome_kw and not local:
        assert any(o.maxed for o in outs) and max(o.counters["ZI"] for o in outs if o.counters) > 50


def test_random_1_to_n_is_a_permutation_in_every_mode():
    """seen-list mode (n >= 128), its conversion to the swap list, and the small-set swap list"""
    for n in (1, 2, 5, 127, 128, 200, 1000):
        r = Random1toN()
        r.init(n, False)
        rnd = RandomSource(n)
        got = []
        while not r.done():
            got.append(r.next(rnd))
        assert sorted(got) == list(range(n)), n


# ---------------------------------------------------------------------------------------------------------- pairs
def _run_paired_engine(index, reads, quals, names, ref_names, preset, local=False, off_size=4):
    """reads / quals / names interleaved (mate 1, mate 2, ...) -> (SAM lines, per-pair results)"""
    from bowtie2_b200.lib import PAIR_RESULT
    from bowtie2_b200.policy_engine import PairedPolicyEngine
    eng = PairedPolicyEngine(OracleBackend(Oracle(index), off_size=off_size, local=local), preset, local=local)
    n = len(reads)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, max(len(r) for r in reads) + 64), dtype=np.uint8)
    pairs = np.zeros(n // 2, dtype=PAIR_RESULT)
    outs = []
    for i in range(n // 2):
        pr = eng.align_pair(reads[2 * i], quals[2 * i], names[2 * i], reads[2 * i + 1], quals[2 * i + 1], names[2 * i + 1])
        outs.append(pr)
        pairs[i]["pair_type"] = pr.pair_type
        for k in range(2):
            r, j = pr.mates[k], 2 * i + k
            if r.aligned:
                _fill(res, ops, j, r, reads[j])
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ref_names, read_names=names, pairs=pairs)
    return lines.rstrip("\n").split("\n"), outs


@pytest.mark.parametrize("fixture,index,ref_names", [
    ("lambda", "lambda_index", ["gi|9626243|ref|NC_001416.1|"]),
    ("rep", "rep_index", ["ctg1", "ctg2"]),
])
def test_paired_sam_identical_to_golden(fixture, index, ref_names, request):
    """concordant, discordant and unpaired-mate records, the locus among equal repeats, MAPQ, XS:i, YS:i, TLEN: every line"""
    base = request.getfixturevalue(index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, f"{fixture}_P_sensitive.sam")) if not l.startswith("@")]
    npairs = len(golden) // 2
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_1.fq"), npairs)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_2.fq"), npairs)
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    lines, outs = _run_paired_engine(base, il(r1, r2), il(q1, q2), il(n1, n2), ref_names, "sensitive")
    assert lines == golden
    if fixture == "rep":
        types = np.bincount([o.pair_type for o in outs], minlength=4)
        assert types[1] > 300 and types[2] > 20 and types[3] > 10
        assert sum(o.n_concord >= 2 for o in outs) > 20


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("preset,rdlen,n,ins_sd,genome_kw", [
    # configuration 5: .bt2l index (64-bit offsets: the RNG draws of eeSaTups widen to 64 bits), 2x150, --sensitive
    ("large:sensitive", 150, 150, 120, dict(contig_len=120000, repeat_frac=0.6, repeat_len=250, repeat_copies=400)),
    ("local:sensitive", 100, 200, 60, {}),
    ("local:very-sensitive", 150, 100, 120, dict(contig_len=120000, repeat_frac=0.6, repeat_len=250, repeat_copies=400)),
    ("sensitive", 100, 250, 60, {}),
    ("very-sensitive", 150, 200, 60, {}),                  # the headline configuration's preset and read length
    ("fast", 50, 250, 150, {}),
    ("sensitive", 100, 150, 120, dict(contig_len=120000, repeat_frac=0.6, repeat_len=250, repeat_copies=400)),
])
def test_paired_sam_and_work_counters_identical_to_reference_program(tmp_path, preset, rdlen, n, ins_sd, genome_kw):
    local = preset.startswith("local:")
    large = preset.startswith("large:")
    sfx = "l" if large else "s"
    preset = preset.split(":")[-1]
    kw = dict(n_contigs=3, contig_len=40000, seed=11, repeat_frac=0.05, repeat_len=300, repeat_copies=12, n_gap=37)
    kw.update(genome_kw)
    genome = synth.make_genome(**kw)
    fa, base = str(tmp_path / "g.fa"), str(tmp_path / "g")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-" + sfx), "--seed", "0", "--quiet", fa, base])
    reads, quals, _ = synth.make_pairs(genome, n, rdlen, seed=7 + rdlen, sub_rate=0.02, indel_rate=0.003, hard_frac=0.2, hard_period=12,
                                       ins_sd=ins_sd)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, reads[0::2], quals[0::2])
    synth.write_fastq(f2, reads[1::2], quals[1::2])
    out = subprocess.check_output([ref_bin("bowtie2-align-" + sfx), *(["--local"] if local else []), "--" + preset, "--seed", "0", "-p", "1",
                                   "--read-times", "-x", base, "-1", f1, "-2", f2], stderr=subprocess.DEVNULL).decode()
    full = [l for l in out.split("\n") if l and not l.startswith("@")]
    want = ["\t".join(l.split("\t")[:11] + [x for x in l.split("\t")[11:] if x[:2] in KEEP]) for l in full]
    names = [f"r{i // 2}" for i in range(2 * n)]
    lines, outs = _run_paired_engine(base, reads, quals, names, [f"chr{k + 1}" for k in range(len(genome))], preset, local, 8 if large else 4)
    bad = [i for i in range(n) if lines[2 * i:2 * i + 2] != want[2 * i:2 * i + 2]]
    assert not bad, (len(bad), lines[2 * bad[0]], want[2 * bad[0]])
    for i, o in enumerate(outs):
        tags = {t[:2]: t[5:] for t in full[2 * i].split("\t")[11:]}
        if "ZI" in tags:
            c = o.counters
            assert (c["ZI"], c["XD"], c["XU"], c["YR"]) == tuple(int(tags[k]) for k in ("ZI", "XD", "XU", "YR")), (i, c)


# ------------------------------------------------------------------------------------------------- options
def _synth_index(tmp_path):
    genome = synth.make_genome(n_contigs=3, contig_len=40000, seed=11, repeat_frac=0.15, repeat_len=300, repeat_copies=12, n_gap=37)
    fa, base = str(tmp_path / "g.fa"), str(tmp_path / "g")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", fa, base])
    return genome, base


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,kw", [
    (["--nofw"], dict(nofw=True)),
    (["--norc"], dict(norc=True)),
    (["--seed", "7"], dict(seed=7)),
    (["-D", "5", "-R", "1", "-L", "20"], dict(dp_fail_streak=5, seed_rounds=1, seed_len=20)),
])
def test_unpaired_options(tmp_path, args, kw):
    genome, base = _synth_index(tmp_path)
    reads, quals, _ = synth.make_reads(genome, 300, 100, seed=77, sub_rate=0.02, indel_rate=0.003)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, reads, quals)
    cmd = [ref_bin("bowtie2-align-s"), "--sensitive", "-p", "1", "-x", base, "-U", fq] + args
    if "--seed" not in args:
        cmd += ["--seed", "0"]
    out = subprocess.check_output(cmd, stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    O = Oracle(base)
    eng = PolicyEngine(OracleBackend(O), "sensitive", **kw)
    n = len(reads)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, 164), dtype=np.uint8)
    for i in range(n):
        r = eng.align_read(reads[i], quals[i], f"r{i}")
        if r.aligned:
            _fill(res, ops, i, r, reads[i])
    names = [f"r{i}" for i in range(n)]
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ["chr1", "chr2", "chr3"], read_names=names).rstrip("\n").split("\n")
    assert lines == want


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,pe_kw,eng_kw", [
    (["--no-mixed"], {}, dict(mixed=False)),
    (["--no-discordant"], {}, dict(discord=False)),
    (["--no-mixed", "--no-discordant"], {}, dict(mixed=False, discord=False)),
    (["-I", "250", "-X", "380"], dict(minfrag=250, maxfrag=380), {}),
    (["--ff"], dict(pol=1), {}),
    (["--rf"], dict(pol=4), {}),
    (["--no-contain", "--no-overlap"], dict(contain_ok=False, olap_ok=False), {}),
    (["--dovetail"], dict(dovetail_ok=True), {}),
    (["--nofw"], {}, dict(nofw=True)),
])
def test_paired_options(tmp_path, args, pe_kw, eng_kw):
    from bowtie2_b200 import policy
    from bowtie2_b200.lib import PAIR_RESULT
    from bowtie2_b200.policy_engine import PairedPolicyEngine
    genome, base = _synth_index(tmp_path)
    n = 200
    reads, quals, _ = synth.make_pairs(genome, n, 100, seed=31, sub_rate=0.02, indel_rate=0.003, hard_frac=0.2, hard_period=12,
                                       ins_mean=300, ins_sd=90)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, reads[0::2], quals[0::2])
    synth.write_fastq(f2, reads[1::2], quals[1::2])
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-1", f1, "-2", f2] + args,
                                  stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    eng = PairedPolicyEngine(OracleBackend(Oracle(base)), "sensitive", pe=policy.PairedEndPolicy(**pe_kw), **eng_kw)
    res = np.zeros(2 * n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((2 * n, 164), dtype=np.uint8)
    pairs = np.zeros(n, dtype=PAIR_RESULT)
    names = [f"r{i // 2}" for i in range(2 * n)]
    for i in range(n):
        pr = eng.align_pair(reads[2 * i], quals[2 * i], names[2 * i], reads[2 * i + 1], quals[2 * i + 1], names[2 * i + 1])
        pairs[i]["pair_type"] = pr.pair_type
        for k in range(2):
            if pr.mates[k].aligned:
                _fill(res, ops, 2 * i + k, pr.mates[k], reads[2 * i + k])
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ["chr1", "chr2", "chr3"], read_names=names,
                       pairs=pairs).rstrip("\n").split("\n")
    bad = [i for i in range(n) if lines[2 * i:2 * i + 2] != want[2 * i:2 * i + 2]]
    assert not bad, (len(bad), lines[2 * bad[0]:2 * bad[0] + 2], want[2 * bad[0]:2 * bad[0] + 2])


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("local", [False, True])
def test_edge_case_reads(tmp_path, local):
    """1-200 bp reads, Ns (a few / all), random reads, extreme qualities, reads shorter than the seed: the filters (YF:Z:LN /
    NS / SC), the short-read seed handling and the ungapped path, against the reference program."""
    genome = synth.make_genome(n_contigs=3, contig_len=30000, seed=5, repeat_frac=0.1, repeat_len=200, repeat_copies=10, n_gap=60)
    fa, base, fq = str(tmp_path / "g.fa"), str(tmp_path / "g"), str(tmp_path / "r.fq")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", fa, base])
    rng = np.random.default_rng(1)
    reads, quals = [], []
    n = 400
    for i in range(n):
        ln = int(rng.choice([1, 2, 3, 5, 10, 15, 19, 20, 21, 22, 23, 25, 30, 33, 40, 49, 54, 64, 99, 104, 109, 149, 200]))
        c = int(rng.integers(0, 3))
        p = int(rng.integers(0, len(genome[c]) - ln))
        r = genome[c][p:p + ln].copy()
        if rng.random() < 0.5:
            r = np.array([4 if x > 3 else 3 - x for x in r[::-1]], dtype=np.uint8)
        k = rng.random()
        if k < 0.2:
            for _ in range(int(rng.integers(1, 4))):
                r[int(rng.integers(0, ln))] = 4
        elif k < 0.25:
            r[:] = 4
        elif k < 0.6:
            for _ in range(int(rng.integers(1, 5))):
                j = int(rng.integers(0, ln))
                r[j] = (r[j] + 1) % 4 if r[j] < 4 else r[j]
        elif k < 0.65:
            r = rng.integers(0, 4, ln).astype(np.uint8)
        reads.append(r)
        quals.append(rng.choice([35, 43, 53, 63, 73, 74], ln).astype(np.uint8))
    synth.write_fastq(fq, reads, quals)
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), *(["--local"] if local else []), "--sensitive", "--seed", "0", "-p", "1",
                                   "-x", base, "-U", fq], stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    eng = PolicyEngine(OracleBackend(Oracle(base), local=local), "sensitive", local=local)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, 264), dtype=np.uint8)
    filt = set()
    for i in range(n):
        r = eng.align_read(reads[i], quals[i], f"r{i}")
        filt.add(r.filtered)
        if r.aligned:
            _fill(res, ops, i, r, reads[i])
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ["chr1", "chr2", "chr3"],
                       read_names=[f"r{i}" for i in range(n)], local=local).rstrip("\n").split("\n")
    bad = [i for i in range(n) if lines[i] != want[i]]
    assert not bad, (len(bad), lines[bad[0]], want[bad[0]])
    assert {"LN", "NS"}.issubset(filt) and (("SC" in filt) == local)


# --------------------------------------------------------------------------------------------------- waves
def test_wave_scheduler_gives_the_sequential_answers(lambda_index, rep_index):
    """many reads advanced together, one batched backend call per primitive and wave: same results as read by read"""
    from bowtie2_b200.policy_engine import PairedPolicyEngine
    from bowtie2_b200.policy_waves import ItemwiseBatch, WaveScheduler
    names, reads, quals = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), 400)
    O = Oracle(lambda_index)
    seq = [PolicyEngine(OracleBackend(O), "sensitive").align_read(reads[i], quals[i], names[i]) for i in range(len(reads))]
    ws = WaveScheduler(ItemwiseBatch(OracleBackend(O)), lambda: PolicyEngine(None, "sensitive"), max_inflight=128)
    got = ws.run_reads(reads, quals, names)
    key = lambda r: (r.aligned, r.filtered, r.xs, r.mapq, None if r.aln is None else (r.aln.tidx, r.aln.refoff, r.aln.fw, r.aln.score, tuple(r.aln.edits)))
    assert [key(r) for r in got] == [key(r) for r in seq]
    # each wave answers every blocked read: far fewer backend calls than requests
    assert ws.n_waves < 100 and sum(ws.n_calls.values()) * 10 < sum(ws.n_requests.values())
    # pairs, on the repeat-rich set
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_1.fq"), 150)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_2.fq"), 150)
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    R, Q, N = il(r1, r2), il(q1, q2), il(n1, n2)
    O2 = Oracle(rep_index)
    eng = PairedPolicyEngine(OracleBackend(O2), "sensitive")
    seqp = [eng.align_pair(R[2 * i], Q[2 * i], N[2 * i], R[2 * i + 1], Q[2 * i + 1], N[2 * i + 1]) for i in range(150)]
    ws2 = WaveScheduler(ItemwiseBatch(OracleBackend(O2)), lambda: PairedPolicyEngine(None, "sensitive"))
    gotp = ws2.run_pairs(R, Q, N)
    pkey = lambda p: (p.pair_type, [key(m) for m in p.mates])
    assert [pkey(p) for p in gotp] == [pkey(p) for p in seqp]


@pytest.mark.parametrize("local", [False, True])
def test_gpu_batch_backend_logic_over_a_fake_device(rep_index, local):
    """policy_backend_gpu.GpuBatchBackend (grouping, DP chunking, array conversions) with the oracle answering behind the
    entry-point conventions: waves over it == the engine over the plain oracle backend, for reads and pairs."""
    from bowtie2_b200.policy_backend_gpu import GpuBackend, GpuBatchBackend
    from bowtie2_b200.policy_engine import PairedPolicyEngine
    from bowtie2_b200.policy_waves import WaveScheduler
    from fake_gpu import FakeGpu
    O = Oracle(rep_index)
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_1.fq"), 120)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_2.fq"), 120)
    key = lambda r: (r.aligned, r.filtered, r.xs, r.mapq, None if r.aln is None else (r.aln.tidx, r.aln.refoff, r.aln.fw, r.aln.score,
                                                                                   r.aln.trim5, r.aln.trim3, r.aln.refns, tuple(r.aln.edits)))
    ref_eng = PolicyEngine(OracleBackend(O, local=local), "sensitive", local=local)
    want = [key(ref_eng.align_read(r1[i], q1[i], n1[i])) for i in range(len(r1))]
    fake = FakeGpu(O)
    bb = GpuBatchBackend(fake, local)
    bb.DP_CHUNK = 16                                              # several chunks per wave
    ws = WaveScheduler(bb, lambda: PolicyEngine(None, "sensitive", local=local), max_inflight=64)
    assert [key(r) for r in ws.run_reads(r1, q1, n1)] == want
    assert fake.calls < 400                                        # batched: far fewer entry-point calls than the ~2000 requests
    # per-item view
    eng1 = PolicyEngine(GpuBackend(FakeGpu(O), local), "sensitive", local=local)
    assert [key(eng1.align_read(r1[i], q1[i], n1[i])) for i in range(30)] == want[:30]
    # pairs
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    R, Q, N = il(r1, r2), il(q1, q2), il(n1, n2)
    pe = PairedPolicyEngine(OracleBackend(O, local=local), "sensitive", local=local)
    wantp = [(p.pair_type, [key(m) for m in p.mates]) for p in
             (pe.align_pair(R[2 * i], Q[2 * i], N[2 * i], R[2 * i + 1], Q[2 * i + 1], N[2 * i + 1]) for i in range(len(r1)))]
    ws2 = WaveScheduler(GpuBatchBackend(FakeGpu(O), local), lambda: PairedPolicyEngine(None, "sensitive", local=local))
    assert [(p.pair_type, [key(m) for m in p.mates]) for p in ws2.run_pairs(R, Q, N)] == wantp


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_sam_output_options(tmp_path):
    """--xeq, --no-unal, --rg-id / --rg: records and header against the reference program"""
    from bowtie2_b200.lib import IndexFile, sam_header
    genome, base = _synth_index(tmp_path)
    reads, quals, _ = synth.make_reads(genome, 300, 100, seed=78, sub_rate=0.02, indel_rate=0.004)
    rng = np.random.default_rng(3)
    for i in range(0, 300, 9):
        reads[i] = rng.integers(0, 4, 100).astype(np.uint8)          # unalignable
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, reads, quals)
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-U", fq, "--xeq",
                                   "--no-unal", "--rg-id", "grp1", "--rg", "SM:sample7", "--rg", "PL:synthetic"],
                                  stderr=subprocess.DEVNULL).decode()
    want_hdr = [l for l in out.split("\n") if l.startswith("@") and not l.startswith("@PG")]
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    eng = PolicyEngine(OracleBackend(Oracle(base)), "sensitive")
    n = len(reads)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, 164), dtype=np.uint8)
    for i in range(n):
        r = eng.align_read(reads[i], quals[i], f"r{i}")
        if r.aligned:
            _fill(res, ops, i, r, reads[i])
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ["chr1", "chr2", "chr3"],
                       read_names=[f"r{i}" for i in range(n)], xeq=True, no_unal=True, rg_id="grp1", threads=3).rstrip("\n").split("\n")
    assert lines == want and len(want) < n and any("X" in l.split("\t")[5] for l in want) and any("D" in l.split("\t")[5] or "I" in l.split("\t")[5] for l in want)
    f = IndexFile(base)
    hdr = sam_header(load_library(), f.ref_names, f.ref_lens, rg_id="grp1", rg_fields=["SM:sample7", "PL:synthetic"])
    assert hdr.rstrip("\n").split("\n") == want_hdr


# ------------------------------------------------------------------------------------------------ -k / -a
def _multi_sam(outs, reads, quals, names, ref_names, local=False):
    """SAM for results that carry secondary alignments: one formatter record per reported alignment (the read repeated),
    secondaries marked in found's bit 8 (FLAG 256) with MAPQ 255; all records of a read share its XS:i"""
    R, Q, N, rows = [], [], [], []
    for i, r in enumerate(outs):
        alns = [None] if not r.aligned else [r.aln] + (r.secondary or [])
        for j, a in enumerate(alns):
            R.append(reads[i]); Q.append(quals[i]); N.append(names[i])
            rows.append((r, a, j > 0))
    if not R:
        return []
    res = np.zeros(len(R), dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((len(R), max(len(x) for x in R) + 64), dtype=np.uint8)
    from bowtie2_b200.policy_engine import ReadResult
    for j, (r, a, sec) in enumerate(rows):
        if a is None:
            continue
        _fill(res, ops, j, ReadResult(aligned=True, aln=a, xs=r.xs, mapq=255 if sec else r.mapq), R[j])
        if sec:
            res[j]["found"] |= 0x100
    return sam_format(load_library(), ReadBatch.from_list(R, Q), res, ops, ref_names, read_names=N, local=local).rstrip("\n").split("\n")


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,kw", [(["-k", "3"], dict(k=3)), (["-k", "12"], dict(k=12)), (["-a"], dict(all_hits=True)), (["-k", "1"], dict(k=1))])
def test_k_and_all_modes_unpaired(tmp_path, args, kw):
    genome, base = _synth_index(tmp_path)
    reads, quals, _ = synth.make_reads(genome, 250, 100, seed=79, sub_rate=0.02, indel_rate=0.003)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, reads, quals)
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-U", fq] + args,
                                  stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    eng = PolicyEngine(OracleBackend(Oracle(base)), "sensitive", **kw)
    names = [f"r{i}" for i in range(len(reads))]
    outs = [eng.align_read(reads[i], quals[i], names[i]) for i in range(len(reads))]
    lines = _multi_sam(outs, reads, quals, names, ["chr1", "chr2", "chr3"])
    assert lines == want, next((a, b) for a, b in zip(lines, want) if a != b)
    if kw.get("k", 2) > 1:
        assert sum(int(l.split("\t")[1]) & 256 != 0 for l in want) > 20


def _multi_sam_pairs(outs, reads, quals, names, ref_names, local=False):
    """records of pairs that carry secondary alignments, in the reference's order (AlnSink::reportHits, aln_sink.h:640-735):
    concordant pairs one after the other; otherwise every record of mate 1, then every record of mate 2 (each printed
    with the opposite mate's primary as its mate), an unaligned mate's record last.  Every record comes from a formatter pair entry; `keep` picks its lines."""
    from bowtie2_b200.lib import PAIR_RESULT
    from bowtie2_b200.policy_engine import ReadResult
    R, Q, N, ent, keep = [], [], [], [], []
    for i, pr in enumerate(outs):
        m1, m2 = pr.mates
        rq = (reads[2 * i], reads[2 * i + 1], quals[2 * i], quals[2 * i + 1], names[2 * i], names[2 * i + 1])

        def add(a1, a2, sec1, sec2, lines):
            R.extend(rq[0:2]); Q.extend(rq[2:4]); N.extend(rq[4:6])
            ent.append((pr, a1, a2, sec1, sec2))
            keep.append(lines)
        a1 = m1.aln if m1.aligned else None
        a2 = m2.aln if m2.aligned else None
        if pr.pair_type == 1:
            add(a1, a2, False, False, (0, 1))
            for (b1, b2) in pr.secondary_pairs or []:
                add(b1, b2, True, True, (0, 1))
        elif not (m1.secondary or m2.secondary):
            add(a1, a2, False, False, (0, 1))
        else:
            # every record of mate 1, then every record of mate 2, an unaligned mate's record last (AlnSinkWrap::finishRead)
            if a1 is not None:
                add(a1, a2, False, False, (0,))
                for b1 in (m1.secondary or []):
                    add(b1, a2, True, False, (0,))
            if a2 is not None:
                add(a1, a2, False, False, (1,))
                for b2 in (m2.secondary or []):
                    add(a1, b2, False, True, (1,))
            if a1 is None:
                add(a1, a2, False, False, (0,))
            if a2 is None:
                add(a1, a2, False, False, (1,))
    n = len(R)
    if n == 0:
        return []
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    ops = np.zeros((n, max(len(x) for x in R) + 64), dtype=np.uint8)
    pairs = np.zeros(n // 2, dtype=PAIR_RESULT)
    for e, (pr, a1, a2, sec1, sec2) in enumerate(ent):
        pairs[e]["pair_type"] = pr.pair_type
        for k, (a, sec, m) in enumerate(((a1, sec1, pr.mates[0]), (a2, sec2, pr.mates[1]))):
            if a is None:
                continue
            _fill(res, ops, 2 * e + k, ReadResult(aligned=True, aln=a, xs=m.xs, mapq=255 if sec else m.mapq), R[2 * e + k])
            if sec:
                res[2 * e + k]["found"] |= 0x100
    lines = sam_format(load_library(), ReadBatch.from_list(R, Q), res, ops, ref_names, read_names=N, pairs=pairs, local=local).rstrip("\n").split("\n")
    out = []
    for e, k in enumerate(keep):
        pair_lines = lines[2 * e:2 * e + 2]
        by_mate = {bool(int(l.split("\t")[1]) & 128): l for l in pair_lines}
        if k == (0, 1):
            out.extend(pair_lines)
        else:
            out.append(by_mate[k[0] == 1])
    return out


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,kw", [(["-k", "3"], dict(k=3)), (["-a"], dict(all_hits=True))])
def test_k_and_all_modes_paired(tmp_path, args, kw):
    from bowtie2_b200.policy_engine import PairedPolicyEngine
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
    eng = PairedPolicyEngine(OracleBackend(Oracle(base)), "sensitive", **kw)
    names = [f"r{i // 2}" for i in range(2 * n)]
    outs = [eng.align_pair(reads[2 * i], quals[2 * i], names[2 * i], reads[2 * i + 1], quals[2 * i + 1], names[2 * i + 1]) for i in range(n)]
    lines = _multi_sam_pairs(outs, reads, quals, names, ["chr1", "chr2", "chr3"])
    diff = [(a, b) for a, b in zip(lines, want) if a != b]
    assert len(lines) == len(want) and not diff, (len(lines), len(want), diff[:1])


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("local", [False, True])
def test_edge_case_pairs(tmp_path, local):
    """mates of 1-60 bp (shorter than the seed, filtered by length / Ns / minimum score), fragments barely longer than a mate"""
    from bowtie2_b200.policy_engine import PairedPolicyEngine
    genome = synth.make_genome(n_contigs=3, contig_len=30000, seed=5, repeat_frac=0.1, repeat_len=200, repeat_copies=10, n_gap=60)
    fa, base = str(tmp_path / "g.fa"), str(tmp_path / "g")
    synth.write_fasta(fa, genome)
    subprocess.check_call([ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", fa, base])
    rng = np.random.default_rng(3)
    R, Q = [], []
    n = 200
    for i in range(n):
        c = int(rng.integers(0, 3)); frag = int(rng.integers(60, 400)); st = int(rng.integers(0, len(genome[c]) - frag))
        l1 = min(int(rng.choice([1, 5, 12, 18, 21, 22, 25, 40, 60])), frag)
        l2 = min(int(rng.choice([1, 8, 15, 20, 23, 30, 50, 60])), frag)
        f = genome[c][st:st + frag]
        a = f[:l1].copy()
        b = np.array([4 if x > 3 else 3 - x for x in f[-l2:][::-1]], dtype=np.uint8)
        if rng.random() < 0.3 and l1 > 3:
            a[int(rng.integers(0, l1))] = 4
        if rng.random() < 0.5:
            a, b = b, a
        R += [a, b]
        Q += [np.full(len(a), 73, np.uint8), np.full(len(b), 53, np.uint8)]
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, R[0::2], Q[0::2])
    synth.write_fastq(f2, R[1::2], Q[1::2])
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), *(["--local"] if local else []), "--sensitive", "--seed", "0", "-p", "1",
                                   "-x", base, "-1", f1, "-2", f2], stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    eng = PairedPolicyEngine(OracleBackend(Oracle(base), local=local), "sensitive", local=local)
    names = [f"r{i // 2}" for i in range(2 * n)]
    outs = [eng.align_pair(R[2 * i], Q[2 * i], names[2 * i], R[2 * i + 1], Q[2 * i + 1], names[2 * i + 1]) for i in range(n)]
    lines = _multi_sam_pairs(outs, R, Q, names, ["chr1", "chr2", "chr3"], local=local)
    bad = [i for i in range(n) if lines[2 * i:2 * i + 2] != want[2 * i:2 * i + 2]]
    assert not bad, (len(bad), lines[2 * bad[0]:2 * bad[0] + 2], want[2 * bad[0]:2 * bad[0] + 2])
