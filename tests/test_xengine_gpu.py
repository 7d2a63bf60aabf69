"""GPU: the exact search policy as a device-side state machine in waves (bt2g_xengine_*, csrc/xengine.cuh + xengine.cu).
Its SAM must be byte-identical to the reference program's: the committed golden files (made by the reference itself) and fresh
runs of the reference binary (oracle/_ref) on synthetic repeat-rich genomes, paired and unpaired, end-to-end and --local,
.bt2 and .bt2l.  All through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bowtie2_b200 import Bt2Gpu
    return Bt2Gpu(0)


def _il(a, b):
    return [x for p in zip(a, b) for x in p]


def _run(g, reads, quals, names, preset, paired, local, ref_names, accel=False, **kw):
    from bowtie2_b200.lib import ReadBatch, XEngine, load_library, policy_params, sam_format
    batch = ReadBatch.from_list(reads, quals)
    if accel:
        g.build_dense_sa(0)
    prm = policy_params(preset, local=local, paired=paired, **kw)
    units = len(reads) // (2 if paired else 1)
    eng = XEngine(g, prm, units, max(len(r) for r in reads))
    try:
        res, ops, pairs, stats = eng.align(batch, names)
        res2, ops2, pairs2, _ = eng.align(batch, names)                      # idempotent (state is reset per batch)
        assert res.tobytes() == res2.tobytes()
    finally:
        eng.close()
        if accel:
            g.build_dense_sa(-1)
    lines = sam_format(load_library(), batch, res, ops, ref_names, read_names=names, pairs=pairs, local=local).rstrip("\n").split("\n")
    return lines, stats


@pytest.mark.parametrize("fixture,index,ref_names,local,paired", [
    ("lambda_U_sensitive", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], False, False),
    ("lambda_U_local", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], True, False),
    ("lambda_P_sensitive", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], False, True),
    ("rep_U_sensitive", "rep_index", ["ctg1", "ctg2"], False, False),
    ("rep_P_sensitive", "rep_index", ["ctg1", "ctg2"], False, True),
])
@pytest.mark.parametrize("accel", [False, True])
def test_device_engine_sam_identical_to_golden(fixture, index, ref_names, local, paired, accel, request):
    from conftest import GOLDEN, read_fastq_codes
    g = _gpu()
    g.load_index_files(request.getfixturevalue(index))
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, fixture + ".sam")) if not l.startswith("@")]
    pre = fixture.split("_")[0]
    n = len(golden) // (2 if paired else 1)
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, pre + "_reads_1.fq"), n)
    if paired:
        n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, pre + "_reads_2.fq"), n)
        R, Q, N = _il(r1, r2), _il(q1, q2), _il(n1, n2)
    else:
        R, Q, N = r1, q1, n1
    lines, stats = _run(g, R, Q, N, "sensitive", paired, local, ref_names, accel)
    bad = [i for i in range(len(golden)) if lines[i] != golden[i]]
    assert not bad, (len(bad), lines[bad[0]], golden[bad[0]], stats)
    # (local mode on these noisy reads yields candidate lists beyond the device engine's per-problem capacity for some reads:
    # those units are answered by the coroutine engine, same records)
    assert stats["fallback_units"] * (5 if local else 20) <= max(n, 20), stats
    g.close()


@pytest.mark.parametrize("paired,local,large,preset,rdlen", [
    (True, False, False, "very-sensitive", 150),              # configs[2]'s preset and read length
    (False, False, False, "sensitive", 100),                  # configs[1]
    (False, True, False, "very-sensitive", 250),              # configs[3]'s mode (local; i16 territory)
    (True, False, True, "sensitive", 150),                    # configs[4]: .bt2l, 64-bit RNG draws
    (True, True, False, "sensitive", 100),
])
def test_device_engine_equals_the_reference_program(tmp_path, paired, local, large, preset, rdlen):
    """repeat-rich synthetic genome (families of 150 copies, an N gap), reads with substitutions and indels: every SAM record
    equal to the reference program's (run here, -p 1 --reorder)"""
    from bowtie2_b200 import synth
    from oracle_lib import have_reference, ref_bin
    if not have_reference():
        pytest.skip("oracle/_ref not built")
    g = _gpu()
    genome = synth.make_genome(n_contigs=3, contig_len=200000, seed=17, repeat_frac=0.2, repeat_len=400, repeat_copies=150, n_gap=53)
    fa, base = str(tmp_path / "g.fa"), str(tmp_path / "g")
    synth.write_fasta(fa, genome)
    sfx = "l" if large else "s"
    subprocess.check_call([ref_bin("bowtie2-build-" + sfx), "--seed", "0", "--quiet", fa, base])
    n = 3000
    if paired:
        reads, quals, _ = synth.make_pairs(genome, n, rdlen, seed=43, sub_rate=0.01, indel_rate=0.001, ins_mean=350, ins_sd=40)
        names = [f"r{i // 2}" for i in range(2 * n)]
        f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
        synth.write_fastq(f1, reads[0::2], quals[0::2]); synth.write_fastq(f2, reads[1::2], quals[1::2])
        io = ["-1", f1, "-2", f2]
    else:
        reads, quals, _ = synth.make_reads(genome, n, rdlen, seed=43, sub_rate=0.01, indel_rate=0.001)
        names = [f"r{i}" for i in range(n)]
        fq = str(tmp_path / "r.fq")
        synth.write_fastq(fq, reads, quals)
        io = ["-U", fq]
    out = subprocess.check_output([ref_bin("bowtie2-align-" + sfx), *(["--local"] if local else []), "--" + preset + ("-local" if local else ""),
                                   "--seed", "0", "-p", "4", "--reorder", "-x", base] + io, stderr=subprocess.DEVNULL).decode()
    want = [l for l in out.split("\n") if l and not l.startswith("@")]
    ref_names = [l.split("\t")[1][3:] for l in out.split("\n") if l.startswith("@SQ")]
    g.load_index_files(base)
    lines, stats = _run(g, reads, quals, None, preset, paired, local, ref_names)      # names=None: "r<index>" made on the device
    bad = [i for i in range(len(want)) if lines[i] != want[i]]
    assert not bad, (len(bad), lines[bad[0]], want[bad[0]], stats)
    if not local:                                            # (local candidate lists often exceed the device engine's capacity: coroutine engine)
        assert stats["fallback_units"] * 10 <= n, stats
    assert stats["seed_dps"] > 100
    g.close()


def test_text_stream_over_device_engines(request):
    """FASTQ text -> SAM text (bowtie2_b200/stream.py) over TWO device engines sharing one context (own streams, own host threads):
    the golden lambda pairs, cut into uneven batches, must come out as the reference program's SAM, in input order"""
    from conftest import GOLDEN
    from bowtie2_b200.lib import XEngine, policy_params
    from bowtie2_b200.stream import TextAligner
    g = _gpu()
    g.load_index_files(request.getfixturevalue("lambda_index"))
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_P_sensitive.sam")) if not l.startswith("@")]
    n = len(golden) // 2

    def recs(path):
        lines = open(path, "rb").read().split(b"\n")
        return [b"\n".join(lines[4 * i:4 * i + 4]) + b"\n" for i in range(n)]
    r1, r2 = recs(os.path.join(GOLDEN, "lambda_reads_1.fq")), recs(os.path.join(GOLDEN, "lambda_reads_2.fq"))
    cuts = [0, 700, 1000, 1900, n] if n > 1900 else [0, n // 3, n]
    cuts = sorted(set(min(c, n) for c in cuts))
    items = [(b"".join(r1[a:b]), b"".join(r2[a:b])) for a, b in zip(cuts[:-1], cuts[1:])]
    prm = policy_params("sensitive", paired=True)
    engines = [XEngine(g, prm, max(b - a for a, b in zip(cuts[:-1], cuts[1:])), 512) for _ in range(2)]
    try:
        ta = TextAligner(engines, ["gi|9626243|ref|NC_001416.1|"], paired=True, parse_threads=2, format_threads=2, name_stride=64)
        chunks = []
        written = ta.run(iter(items), lambda v: chunks.append(bytes(v)))     # (the view is valid only inside the sink)
    finally:
        for e in engines:
            e.close()
    lines = b"".join(chunks).decode().rstrip("\n").split("\n")
    bad = [i for i in range(len(golden)) if lines[i] != golden[i]]
    assert written == 2 * n and not bad, (written, len(bad), lines[bad[0]] if bad else None)
    g.close()
