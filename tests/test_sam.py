"""bt2g_sam_format (host code): SAM records rebuilt from alignments and compared, whole line, with the committed
output of the reference program (tests/golden/lambda_U_sensitive.sam: bowtie2 --sensitive -U lambda_reads_1.fq).

The alignment fed to the formatter is the oracle DP's (C restatement of SwAligner) for the window around the
reference's reported position, converted from the reference's Edit representation to the device op string;
MAPQ and XS:i depend on the sequential search policy and are taken from the golden record."""
import os

import numpy as np
import pytest

from bowtie2_b200 import policy
from bowtie2_b200.lib import (OP_MATCH, OP_MM, OP_READGAP, OP_REFGAP, READ_RESULT, ReadBatch, load_library, sam_format)
from conftest import GOLDEN, read_fastq_codes
from oracle_lib import Oracle, oracle_dp



def _edits_to_ops(edits, read, fw, trim5=0, trim3=0):
    """Inverse of lib.ops_to_edits: reference Edit list (positions relative to the soft-trimmed extent, from the 5'
    end of the read) -> op string (last aligned read row first), every op that consumes a reference base carrying
    its code."""
    code = {ord(c): i for i, c in enumerate("ACGTN")}
    rdlen = len(read)
    seq = read if fw else np.array([4 if c > 3 else 3 - c for c in read[::-1]], dtype=np.uint8)
    row0 = trim5 if fw else trim3                      # rows soft-trimmed on the left, in reference orientation
    ext = rdlen - trim5 - trim3
    ed = [list(e) for e in edits]
    gap = ord("-")
    if not fw:
        ed = ed[::-1]
        for e in ed:
            e[0] = ext - e[0] - (0 if e[2] == gap else 1)            # read gap: qchr == '-'
    fwd = []
    k = 0
    for rel in range(ext):
        while k < len(ed) and ed[k][0] == rel and ed[k][2] == gap:     # read gaps before this row
            fwd.append(OP_READGAP | (code[ed[k][1]] << 2)); k += 1
        if k < len(ed) and ed[k][0] == rel:
            if ed[k][1] == gap:
                fwd.append(OP_REFGAP)
            else:
                fwd.append(OP_MM | (code[ed[k][1]] << 2))
            k += 1
        else:
            fwd.append(OP_MATCH | (int(seq[row0 + rel]) << 2))
    assert k == len(ed)
    return fwd[::-1]


def test_sam_records_match_reference_program(lambda_index):
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_U_sensitive.sam")) if not l.startswith("@")]
    n = 300
    names, reads, quals = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    O = Oracle(lambda_index)
    sc = policy.Scoring.default(False)
    tlen = 48502
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    maxops = max(len(r) for r in reads) + 64
    ops = np.zeros((n, maxops), dtype=np.uint8)
    use = []
    for i in range(n):
        f = golden[i].split("\t")
        flag = int(f[1])
        if flag & 4:
            use.append(i)
            continue
        tags = {t[:2]: t[5:] for t in f[11:]}
        pos, fw, AS = int(f[3]) - 1, not (flag & 16), int(tags["AS"])
        r = reads[i]
        rdlen = len(r)
        minsc = sc.min_score(rdlen)
        found, rect = policy.frame_seed_extension_rect(pos, rdlen, tlen, sc.max_read_gaps(minsc, rdlen), sc.max_ref_gaps(minsc, rdlen),
                                                       sc.n_ceil(rdlen))
        d = oracle_dp(O, False, r, quals[i], fw, 0, rect, minsc, sc.n_ceil_raw(rdlen), max_alns=8)
        al = [a for a in d["alns"] if a["refoff"] == pos and a["score"] == AS]
        if not al:
            continue                                       # the reference reported an alignment from another window
        a = al[0]
        o = _edits_to_ops(a["edits"], r, fw)
        res[i]["found"] = 2 if (not a["edits"]) else 1
        res[i]["score"] = AS
        res[i]["score2"] = int(tags["XS"]) if "XS" in tags else -(1 << 31)
        res[i]["fw"] = int(fw); res[i]["tidx"] = 0; res[i]["refoff"] = pos; res[i]["nops"] = len(o)
        res[i]["mapq"] = int(f[4]); res[i]["pad"] = int(tags["XN"])
        ops[i, :len(o)] = o
        use.append(i)
    assert len(use) > 0.9 * n
    lib = load_library()
    text = sam_format(lib, ReadBatch.from_list(reads, quals), res, ops, ["gi|9626243|ref|NC_001416.1|"], read_names=names)
    lines = text.rstrip("\n").split("\n")
    assert len(lines) == n
    ngap = 0
    for i in use:
        assert lines[i] == golden[i], (i, lines[i], golden[i])
        ngap += ("I" in golden[i].split("\t")[5]) or ("D" in golden[i].split("\t")[5])
    assert ngap > 3


def test_sam_paired_records_match_reference_program(lambda_index):
    """Paired records (flags, RNEXT/PNEXT/TLEN, YS, YT, mate of an unaligned read) against the golden output for the
    first 200 pairs; alignments of each mate rebuilt through the oracle DP as above."""
    from bowtie2_b200.lib import PAIR_RESULT
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_P_sensitive.sam")) if not l.startswith("@")]
    npairs = len(golden) // 2
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), npairs)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_2.fq"), npairs)
    O = Oracle(lambda_index)
    sc = policy.Scoring.default(False)
    tlen = 48502
    n = 2 * npairs
    reads = [x for p in zip(r1, r2) for x in p]
    quals = [x for p in zip(q1, q2) for x in p]
    names = [x for p in zip(n1, n2) for x in p]
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    maxops = max(len(r) for r in reads) + 64
    ops = np.zeros((n, maxops), dtype=np.uint8)
    pairs = np.zeros(npairs, dtype=PAIR_RESULT)
    # golden lines by (pair, mate): the reference prints the aligned mate first when only mate 2 aligned
    by = {}
    for l in golden:
        f = l.split("\t")
        by[(f[0], 1 if int(f[1]) & 128 else 0)] = l
    ok_pair = []
    for pi in range(npairs):
        good = True
        for mate in range(2):
            i = 2 * pi + mate
            f = by[(names[i], mate)].split("\t")
            flag = int(f[1])
            if flag & 4:
                continue
            tags = {t[:2]: t[5:] for t in f[11:]}
            pos, fw, AS = int(f[3]) - 1, not (flag & 16), int(tags["AS"])
            r = reads[i]
            rdlen = len(r)
            minsc = sc.min_score(rdlen)
            found, rect = policy.frame_seed_extension_rect(pos, rdlen, tlen, sc.max_read_gaps(minsc, rdlen), sc.max_ref_gaps(minsc, rdlen),
                                                           sc.n_ceil(rdlen))
            d = oracle_dp(O, False, r, quals[i], fw, 0, rect, minsc, sc.n_ceil_raw(rdlen), max_alns=8)
            al = [a for a in d["alns"] if a["refoff"] == pos and a["score"] == AS]
            if not al:
                good = False
                continue
            o = _edits_to_ops(al[0]["edits"], r, fw)
            res[i]["found"] = 2 if not al[0]["edits"] else 1
            res[i]["score"] = AS
            res[i]["score2"] = int(tags["XS"]) if "XS" in tags else -(1 << 31)
            res[i]["fw"] = int(fw); res[i]["refoff"] = pos; res[i]["nops"] = len(o)
            res[i]["mapq"] = int(f[4]); res[i]["pad"] = int(tags["XN"])
            ops[i, :len(o)] = o
        yt = by[(names[2 * pi], 0)].split("YT:Z:")[1][:2]
        a1, a2 = res[2 * pi]["found"] != 0, res[2 * pi + 1]["found"] != 0
        pairs[pi]["pair_type"] = 1 if yt == "CP" else (2 if yt == "DP" else (3 if (a1 or a2) else 0))
        if good:
            ok_pair.append(pi)
    assert len(ok_pair) > 0.9 * npairs
    lib = load_library()
    text = sam_format(lib, ReadBatch.from_list(reads, quals), res, ops, ["gi|9626243|ref|NC_001416.1|"], read_names=names, pairs=pairs)
    assert text == sam_format(lib, ReadBatch.from_list(reads, quals), res, ops, ["gi|9626243|ref|NC_001416.1|"], read_names=names, pairs=pairs, threads=5)
    lines = text.rstrip("\n").split("\n")
    assert len(lines) == n
    kinds = set()
    for pi in ok_pair:
        for k in (2 * pi, 2 * pi + 1):
            assert lines[k] == golden[k], (pi, lines[k], golden[k])
            kinds.add(int(golden[k].split("\t")[1]))
    assert {99, 147, 83, 163}.issubset(kinds) and (69 in kinds or 137 in kinds)


def test_sam_local_records_match_reference_program(lambda_index):
    """--local --sensitive-local: soft-clipped records (S in CIGAR, POS at the first aligned base)."""
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_U_local.sam")) if not l.startswith("@")]
    n = len(golden)
    names, reads, quals = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    O = Oracle(lambda_index)
    sc = policy.Scoring.default(True)
    tlen = 48502
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    maxops = max(len(r) for r in reads) + 64
    ops = np.zeros((n, maxops), dtype=np.uint8)
    use, nclip = [], 0
    for i in range(n):
        f = golden[i].split("\t")
        flag = int(f[1])
        if flag & 4:
            use.append(i)
            continue
        tags = {t[:2]: t[5:] for t in f[11:]}
        pos, fw, AS, cig = int(f[3]) - 1, not (flag & 16), int(tags["AS"]), f[5]
        lead = int(cig[:cig.index("S")]) if "S" in cig and cig.index("S") < cig.index("M") else 0
        r = reads[i]
        rdlen = len(r)
        minsc = sc.min_score(rdlen)
        found, rect = policy.frame_seed_extension_rect(pos - lead, rdlen, tlen, sc.max_read_gaps(minsc, rdlen),
                                                       sc.max_ref_gaps(minsc, rdlen), sc.n_ceil(rdlen))
        d = oracle_dp(O, True, r, quals[i], fw, 0, rect, minsc, sc.n_ceil_raw(rdlen), max_cands=16384, max_alns=16)
        al = [a for a in d["alns"] if a["refoff"] == pos and a["score"] == AS]
        if not al:
            continue
        a = al[0]
        o = _edits_to_ops(a["edits"], r, fw, a["trim5"], a["trim3"])
        res[i]["found"] = 1
        res[i]["score"] = AS
        res[i]["score2"] = int(tags["XS"]) if "XS" in tags else -(1 << 31)
        res[i]["fw"] = int(fw); res[i]["refoff"] = pos; res[i]["nops"] = len(o)
        res[i]["trim_left"] = a["trim5"] if fw else a["trim3"]
        res[i]["trim_right"] = a["trim3"] if fw else a["trim5"]
        res[i]["mapq"] = int(f[4]); res[i]["pad"] = int(tags["XN"])
        ops[i, :len(o)] = o
        use.append(i)
        nclip += "S" in cig
    assert len(use) > 0.9 * n and nclip > 30
    text = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ["gi|9626243|ref|NC_001416.1|"], read_names=names)
    lines = text.rstrip("\n").split("\n")
    for i in use:
        assert lines[i] == golden[i], (i, lines[i], golden[i])


def test_fastq_parse_matches_test_reader():
    """bt2g_fastq_parse against the tests' own FASTQ reader on the golden reads, plus chunked parsing and the
    error paths."""
    from bowtie2_b200.lib import fastq_parse
    lib = load_library()
    path = os.path.join(GOLDEN, "lambda_reads_1.fq")
    text = open(path, "rb").read()
    names, reads, quals = read_fastq_codes(path, 10 ** 9)
    batch, nm, used = fastq_parse(lib, text)
    assert used == len(text) and batch.n == len(reads) and nm == names
    for i in (0, 1, 17, len(reads) - 1):
        a, b = int(batch.off[i]), int(batch.off[i + 1])
        assert np.array_equal(batch.seq[a:b], reads[i]) and np.array_equal(batch.qual[a:b], quals[i])
    assert np.array_equal(batch.seq, np.concatenate(reads)) and np.array_equal(batch.qual, np.concatenate(quals))
    # a buffer cut in the middle of a record: whole records only, the rest is left for the next call
    cut = len(text) // 2 + 7
    b1, n1, u1 = fastq_parse(lib, text[:cut])
    b2, n2, u2 = fastq_parse(lib, text[u1:])
    assert u1 <= cut and n1 + n2 == names and b1.n + b2.n == batch.n
    # lower case, '.', IUPAC codes, CRLF
    b3, n3, _ = fastq_parse(lib, b"@x y\r\nacgtN.Rn\r\n+\r\nIIIIIIII\r\n")
    assert n3 == ["x y"] and b3.seq.tolist() == [0, 1, 2, 3, 4, 4, 4, 4]
    with pytest.raises(RuntimeError):
        fastq_parse(lib, b">x\nACGT\n")
    with pytest.raises(RuntimeError):
        fastq_parse(lib, b"@x\nACGT\n+\nII\n@y\nAC\n+\nII\n")


# ---------------------------------------------------------------------------------------------------------------
# repeat-rich fixture (tests/golden/rep_*): XS:i, low MAPQ, two contigs, discordant pairs, mates reported as
# unpaired alignments of a paired read, and the ">1 times" lines of the summary

REP_NAMES, REP_LENS = ["ctg1", "ctg2"], [14000, 10000]


def _rebuild(golden, names, reads, quals, O, paired):
    """Per-read results and op strings for the formatter from the reference's own records: the alignment through the
    oracle DP around the reported position; the policy-dependent fields (MAPQ, XS:i and, for mates of a paired read
    reported as unpaired, whether a second alignment existed) from the record."""
    from bowtie2_b200.lib import PAIR_RESULT, load_library
    lib = load_library()
    sc = policy.Scoring.default(False)
    n = len(reads)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    maxops = max(len(r) for r in reads) + 64
    ops = np.zeros((n, maxops), dtype=np.uint8)
    by = {}
    for l in golden:
        f = l.split("\t")
        by[(f[0], 1 if int(f[1]) & 128 else 0)] = l
    good = np.ones(n, dtype=bool)
    order = []
    for i in range(n):
        l = by[(names[i], (i & 1) if paired else 0)]
        order.append(l)
        f = l.split("\t")
        flag = int(f[1])
        if flag & 4:
            continue
        tags = {t[:2]: t[5:] for t in f[11:]}
        pos, fw, AS = int(f[3]) - 1, not (flag & 16), int(tags["AS"])
        tidx = REP_NAMES.index(f[2])
        r = reads[i]
        rdlen = len(r)
        minsc = sc.min_score(rdlen)
        found, rect = policy.frame_seed_extension_rect(pos, rdlen, REP_LENS[tidx], sc.max_read_gaps(minsc, rdlen),
                                                       sc.max_ref_gaps(minsc, rdlen), sc.n_ceil(rdlen))
        d = oracle_dp(O, False, r, quals[i], fw, tidx, rect, minsc, sc.n_ceil_raw(rdlen), max_alns=8)
        al = [a for a in d["alns"] if a["refoff"] == pos and a["score"] == AS]
        if not al:
            good[i] = False
            continue
        o = _edits_to_ops(al[0]["edits"], r, fw)
        res[i]["found"] = 2 if not al[0]["edits"] else 1
        res[i]["score"] = AS
        res[i]["fw"] = int(fw); res[i]["tidx"] = tidx; res[i]["refoff"] = pos; res[i]["nops"] = len(o)
        res[i]["mapq"] = int(f[4]); res[i]["pad"] = int(tags["XN"])
        if "XS" in tags:
            res[i]["score2"] = int(tags["XS"])
        elif paired and "YT:Z:UP" in l:
            # XS:i is never printed for these; a MAPQ different from the unique-alignment value means a second alignment
            # (discordant mates are unique by definition, and their MAPQ is computed from the pair's score sum)
            uniq = lib.bt2g_mapq(AS, 0, 0, minsc, 0, 1)
            if int(f[4]) != uniq:
                res[i]["score2"] = AS                           # any valid value: only its presence matters
        ops[i, :len(o)] = o
    pairs = None
    if paired:
        pairs = np.zeros(n // 2, dtype=PAIR_RESULT)
        for pi in range(n // 2):
            yt = order[2 * pi].split("YT:Z:")[1][:2]
            a1, a2 = res[2 * pi]["found"] != 0, res[2 * pi + 1]["found"] != 0
            pairs[pi]["pair_type"] = 1 if yt == "CP" else (2 if (a1 and a2) else (3 if (a1 or a2) else 0))
    return res, ops, pairs, good, order


def _load_rep(paired):
    tag = "P" if paired else "U"
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, f"rep_{tag}_sensitive.sam")) if not l.startswith("@")]
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_1.fq"), 10 ** 9)
    if not paired:
        return golden, n1, r1, q1
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_2.fq"), 10 ** 9)
    return (golden, [x for p in zip(n1, n2) for x in p], [x for p in zip(r1, r2) for x in p], [x for p in zip(q1, q2) for x in p])


def test_sam_repeat_genome_unpaired(rep_index):
    golden, names, reads, quals = _load_rep(False)
    res, ops, _, good, order = _rebuild(golden, names, reads, quals, Oracle(rep_index), False)
    assert good.mean() > 0.9
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, REP_NAMES, read_names=names).rstrip("\n").split("\n")
    nxs = nlow = 0
    for i in np.nonzero(good)[0]:
        assert lines[i] == golden[i], (i, lines[i], golden[i])
        nxs += "XS:i:" in golden[i]
        nlow += (not int(golden[i].split("\t")[1]) & 4) and int(golden[i].split("\t")[4]) <= 1
    assert nxs > 80 and nlow > 40


def test_sam_repeat_genome_paired(rep_index):
    golden, names, reads, quals = _load_rep(True)
    res, ops, pairs, good, order = _rebuild(golden, names, reads, quals, Oracle(rep_index), True)
    assert good.mean() > 0.9
    lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, REP_NAMES, read_names=names, pairs=pairs,
                       threads=3).rstrip("\n").split("\n")
    assert len(lines) == len(golden)
    seen = set()
    for pi in range(len(reads) // 2):
        if not (good[2 * pi] and good[2 * pi + 1]):
            continue
        # the reference prints the aligned mate first when only mate 2 aligned: compare by (name, mate)
        got = {bool(int(l.split("\t")[1]) & 128): l for l in lines[2 * pi:2 * pi + 2]}
        for k in (0, 1):
            assert got[bool(k)] == order[2 * pi + k], (pi, got[bool(k)], order[2 * pi + k])
            f = order[2 * pi + k].split("\t")
            seen.add((f[-1] if f[-1].startswith("YT") else [t for t in f if t.startswith("YT")][0], int(f[1])))
        assert lines[2 * pi:2 * pi + 2] == golden[2 * pi:2 * pi + 2]              # and in the same order
    kinds = {k for k, _ in seen}
    flags = {v for _, v in seen}
    assert kinds == {"YT:Z:CP", "YT:Z:DP", "YT:Z:UP"}
    assert {99, 147, 83, 163, 65, 129, 81, 161, 97, 145, 73, 133, 89, 165, 77, 141, 137, 69}.issubset(flags)


@pytest.mark.parametrize("fixture", ["lambda_U_sensitive", "lambda_U_local", "lambda_P_sensitive", "rep_U_sensitive", "rep_P_sensitive"])
def test_alignment_summary_matches_reference_stderr(fixture):
    """bt2g_align_counts_add + bt2g_align_summary against what the reference printed for the same run; the records of
    the run supply found / second-alignment / pair type, as the pipeline's result arrays would."""
    from bowtie2_b200.lib import PAIR_RESULT, align_counts_add, align_summary
    lib = load_library()
    want = open(os.path.join(GOLDEN, fixture + ".summary.txt")).read()
    golden = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN, fixture + ".sam")) if not l.startswith("@")]
    paired = "_P_" in fixture
    local = "local" in fixture
    sc = policy.Scoring.default(local)
    n = len(golden)
    res = np.zeros(n, dtype=READ_RESULT)
    res["score2"] = -(1 << 31)
    # order records as (pair, mate)
    if paired:
        golden = [x for i in range(0, n, 2) for x in sorted(golden[i:i + 2], key=lambda f: int(f[1]) & 128)]
    for i, f in enumerate(golden):
        flag = int(f[1])
        if flag & 4:
            continue
        tags = {t[:2]: t[5:] for t in f[11:]}
        res[i]["found"] = 1
        res[i]["score"] = int(tags["AS"])
        rdlen = len(f[9])
        uniq = lib.bt2g_mapq(int(tags["AS"]), 0, 0, sc.min_score(rdlen), sc.perfect_score(rdlen) if local else 0, 1)
        if "XS" in tags or (int(f[4]) != uniq and paired and "YT:Z:UP" in f):
            res[i]["score2"] = int(tags["AS"])
    pairs = None
    if paired:
        pairs = np.zeros(n // 2, dtype=PAIR_RESULT)
        for pi in range(n // 2):
            yt = [t for t in golden[2 * pi] if t.startswith("YT:Z:")][0][5:]
            a1, a2 = res[2 * pi]["found"] != 0, res[2 * pi + 1]["found"] != 0
            pairs[pi]["pair_type"] = 1 if yt == "CP" else (2 if (a1 and a2) else (3 if (a1 or a2) else 0))
    counts = align_counts_add(lib, None, res, pairs)
    # counters add up over batches
    if n > 10:
        k = (n // 4) * 2
        c2 = align_counts_add(lib, None, res[:k], None if pairs is None else pairs[:k // 2])
        c2 = align_counts_add(lib, c2, res[k:], None if pairs is None else pairs[k // 2:])
        assert c2.tobytes() == counts.tobytes()
    got = align_summary(lib, counts)
    if paired:
        # the second-best concordant PAIR is not part of the pipeline's results (include/bt2g.h): take the reference's
        # split of the concordant pairs into "exactly 1" / ">1", check the total, and compare the rest of the text
        wl = want.split("\n")
        uni1, gt1 = int(wl[3].split()[0]), int(wl[4].split()[0])
        assert int(counts["nconcord_uni1"][0] + counts["nconcord_gt1"][0]) == uni1 + gt1
        counts["nconcord_uni1"], counts["nconcord_gt1"] = uni1, gt1
        got = align_summary(lib, counts)
    assert got == want
    if not paired:
        assert align_summary(lib, counts, discord=False, mixed=False) == want


def test_fastq_parse_multithreaded_equals_serial():
    """bt2g_fastq_parse_mt: pieces cut at record boundaries, parsed concurrently, concatenated in order: same arrays, same
    stopping points under the read limit and with a truncated tail; quality lines starting with '@' do not fool the cutter"""
    from bowtie2_b200.lib import fastq_parse
    lib = load_library()
    rng = np.random.default_rng(4)
    parts = []
    n = 30000
    for i in range(n):
        ln = int(rng.integers(30, 150))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), ln, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        q = bytearray(rng.integers(33, 74, ln).astype(np.uint8).tobytes())
        if i % 7 == 0:
            q[0] = ord("@")                                  # a quality string that looks like a header
        if i % 11 == 0:
            q[0] = ord("+")
        parts.append(b"@read%d some comment\n%s\n+\n%s\n" % (i, seq, bytes(q)))
    text = b"".join(parts)
    assert len(text) > (1 << 21)
    b1, n1, u1 = fastq_parse(lib, text, threads=1)
    for th in (2, 5, 16):
        b2, n2, u2 = fastq_parse(lib, text, threads=th)
        assert u2 == u1 == len(text) and b2.n == b1.n == n
        assert np.array_equal(b1.seq, b2.seq) and np.array_equal(b1.qual, b2.qual) and np.array_equal(b1.off, b2.off)
        assert np.array_equal(n1.rows, n2.rows)
    # read limit: stops at the same record as the serial parser
    b3, n3, u3 = fastq_parse(lib, text, max_reads=12345, threads=1)
    b4, n4, u4 = fastq_parse(lib, text, max_reads=12345, threads=6)
    assert b3.n == b4.n == 12345 and u3 == u4 and np.array_equal(b3.seq, b4.seq) and np.array_equal(n3.rows, n4.rows)
    assert fastq_parse(lib, text[u3:], threads=4)[0].n == n - 12345
    # truncated tail: whole records only
    cut = len(text) - 37
    b5, _, u5 = fastq_parse(lib, text[:cut], threads=1)
    b6, _, u6 = fastq_parse(lib, text[:cut], threads=8)
    assert b5.n == b6.n == n - 1 and u5 == u6 and np.array_equal(b5.off, b6.off)
    # errors surface
    bad = text[:len(text) // 2] + b"@x\nACGT\n+\nII\n" + text[len(text) // 2:]
    with pytest.raises(RuntimeError):
        fastq_parse(lib, bad, threads=4)


def _synthetic_fastq(n, seed, tag, lo=30, hi=150):
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(n):
        ln = int(rng.integers(lo, hi))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), ln, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        q = bytearray(rng.integers(33, 74, ln).astype(np.uint8).tobytes())
        if i % 7 == 0:
            q[0] = ord("@")                                  # a quality string that looks like a header
        parts.append(b"@pair%d/%s a comment\n%s\n+\n%s\n" % (i, tag, seq, bytes(q)))
    return parts


def test_fastq_parse_pairs_equals_two_parses_interleaved():
    """bt2g_fastq_parse_pairs_mt (the two mate texts straight into one interleaved batch) against bt2g_fastq_parse of each text +
    the numpy interleave: same arrays whatever the thread count; reused output buffers carry nothing over; the pair limit, a
    shorter mate file and a truncated tail stop both texts at record boundaries of the same pair"""
    from bowtie2_b200.align import interleave, interleave_names
    from bowtie2_b200.lib import HostBuffers, fastq_parse, fastq_parse_pairs
    lib = load_library()
    n = 26000
    p1, p2 = _synthetic_fastq(n, 5, b"1"), _synthetic_fastq(n, 6, b"2", lo=20, hi=260)
    t1, t2 = b"".join(p1), b"".join(p2)
    assert len(t1) > (1 << 21) and len(t2) > (1 << 21)
    b1, n1, _ = fastq_parse(lib, t1)
    b2, n2, _ = fastq_parse(lib, t2)
    want, want_names = interleave(b1, b2), interleave_names(n1, n2)
    out = HostBuffers()
    for th in (1, 2, 5, 8, 16):
        got, names, u1, u2 = fastq_parse_pairs(lib, t1, t2, threads=th, out=out if th != 2 else None)
        assert u1 == len(t1) and u2 == len(t2) and got.n == 2 * n
        assert np.array_equal(got.off, want.off) and np.array_equal(got.seq, want.seq) and np.array_equal(got.qual, want.qual)
        assert np.array_equal(names.rows, want_names.rows)
    # the same buffers again with a smaller input (and a short first guess of the record count on the way back up)
    k = 1500
    s1, s2 = b"".join(p1[:k]), b"".join(p2[:k])
    got, names, u1, u2 = fastq_parse_pairs(lib, s1, s2, threads=4, out=out)
    assert got.n == 2 * k and u1 == len(s1) and u2 == len(s2)
    assert np.array_equal(got.seq, want.seq[:int(want.off[2 * k])]) and np.array_equal(names.rows, want_names.rows[:2 * k])
    got, names, u1, u2 = fastq_parse_pairs(lib, t1, t2, threads=4, out=out)
    assert got.n == 2 * n and u1 == len(t1) and np.array_equal(got.qual, want.qual) and np.array_equal(names.rows, want_names.rows)
    # pair limit: both texts stop in front of the same pair, the rest parses to the rest
    lim = 12345
    a, an, u1, u2 = fastq_parse_pairs(lib, t1, t2, threads=6, max_pairs=lim)
    assert a.n == 2 * lim and u1 == len(b"".join(p1[:lim])) and u2 == len(b"".join(p2[:lim]))
    b, bn, v1, v2 = fastq_parse_pairs(lib, t1[u1:], t2[u2:], threads=3)
    assert b.n == 2 * (n - lim) and np.array_equal(np.concatenate([a.seq, b.seq]), want.seq)
    assert np.array_equal(np.concatenate([an.rows, bn.rows]), want_names.rows)
    # mate file 2 is shorter / its tail is cut inside a record: whole pairs only, mate 1's cursor waits at the unpaired record
    short2 = b"".join(p2[:n - 3])
    c, _, u1, u2 = fastq_parse_pairs(lib, t1, short2, threads=4)
    assert c.n == 2 * (n - 3) and u2 == len(short2) and u1 == len(b"".join(p1[:n - 3]))
    c, _, u1, u2 = fastq_parse_pairs(lib, t1, t2[:len(t2) - 9], threads=8)
    assert c.n == 2 * (n - 1) and u1 == len(b"".join(p1[:n - 1])) and u2 == len(b"".join(p2[:n - 1]))
    assert np.array_equal(c.seq, want.seq[:int(want.off[2 * (n - 1)])])
    # nothing at all, and small texts (one piece per file)
    e, en, u1, u2 = fastq_parse_pairs(lib, b"", b"", threads=4)
    assert e.n == 0 and u1 == 0 and u2 == 0 and len(en) == 0
    f, fn, _, _ = fastq_parse_pairs(lib, b"@x/1\nacgtN.\n+\nIIIIII\n", b"@x/2\r\nTT\r\n+\r\nII\r\n", threads=4)
    assert f.n == 2 and f.seq.tolist() == [0, 1, 2, 3, 4, 4, 3, 3] and f.off.tolist() == [0, 6, 8] and list(fn) == ["x/1", "x/2"]
    # errors surface from either text
    with pytest.raises(RuntimeError):
        fastq_parse_pairs(lib, t1, t2[:len(t2) // 2] + b"@x\nACGT\n+\nII\n" + t2[len(t2) // 2:], threads=4)
    with pytest.raises(RuntimeError):
        fastq_parse_pairs(lib, b">x\nACGT\n", b"@x\nACGT\n+\nIIII\n")


def test_sam_format_threads_views_and_reused_buffers(rep_index):
    """bt2g_sam_format: the same text from 1 and from several host threads, as bytes and as a view of a reused output buffer, also
    after the buffer held a longer text (the repeat-rich pairs: every pair kind, held mate-1 records included)"""
    from bowtie2_b200.lib import HostBuffers
    golden, names, reads, quals = _load_rep(True)
    res, ops, pairs, good, order = _rebuild(golden, names, reads, quals, Oracle(rep_index), True)
    lib = load_library()
    b = ReadBatch.from_list(reads, quals)
    one = sam_format(lib, b, res, ops, REP_NAMES, read_names=names, pairs=pairs, threads=1, as_bytes=True)
    assert one.decode() == sam_format(lib, b, res, ops, REP_NAMES, read_names=names, pairs=pairs, threads=1)
    out = HostBuffers()
    for th in (2, 3, 7):
        v = sam_format(lib, b, res, ops, REP_NAMES, read_names=names, pairs=pairs, threads=th, as_bytes="view", out=out)
        assert isinstance(v, memoryview) and bytes(v) == one
    k = 200
    part = ReadBatch.from_list(reads[:k], quals[:k])
    v = sam_format(lib, part, res[:k], ops[:k], REP_NAMES, read_names=names[:k], pairs=pairs[:k // 2], threads=2, as_bytes="view", out=out)
    assert bytes(v) == b"".join(x + b"\n" for x in one.split(b"\n")[:k])
    # --no-unal and a read group through the in-place writer
    w1 = sam_format(lib, b, res, ops, REP_NAMES, read_names=names, pairs=pairs, threads=1, no_unal=True, rg_id="grp")
    w4 = sam_format(lib, b, res, ops, REP_NAMES, read_names=names, pairs=pairs, threads=4, no_unal=True, rg_id="grp")
    assert w1 == w4 and all(l.endswith("\tRG:Z:grp") and not int(l.split("\t")[1]) & 4 for l in w1.rstrip("\n").split("\n"))
