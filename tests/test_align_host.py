"""Host plumbing of bowtie2_b200.align (no GPU): FASTQ batching over plain and gzip files, mate interleaving, and the
loud failure of the whole-file entry point when there is no device."""
import gzip
import os
import shutil

import numpy as np
import pytest

from bowtie2_b200.align import FastqStream, align_files, interleave
from conftest import GOLDEN, read_fastq_codes


@pytest.mark.parametrize("gz", [False, True])
def test_fastq_stream_batches(tmp_path, gz):
    src = os.path.join(GOLDEN, "rep_reads_1.fq")
    path = src
    if gz:
        path = str(tmp_path / "r.fq.gz")
        with open(src, "rb") as f, gzip.open(path, "wb") as g:
            shutil.copyfileobj(f, g)
    names, reads, quals = read_fastq_codes(src, 10 ** 9)
    s = FastqStream(path, chunk_bytes=3000)
    got_names, seqs, qs, sizes = [], [], [], []
    while True:
        b, n = s.next_batch(37)
        if b.n == 0:
            break
        sizes.append(b.n)
        got_names += n
        seqs.append(b.seq[:int(b.off[-1])])
        qs.append(b.qual[:int(b.off[-1])])
    s.close()
    assert all(x == 37 for x in sizes[:-1]) and sum(sizes) == len(reads)
    assert got_names == names
    assert np.array_equal(np.concatenate(seqs), np.concatenate(reads)) and np.array_equal(np.concatenate(qs), np.concatenate(quals))


def test_fastq_stream_without_final_newline_and_empty(tmp_path):
    p = tmp_path / "a.fq"
    p.write_text("@x\nACGT\n+\nIIII\n@y\nGG\n+\nII")
    b, n = FastqStream(str(p)).next_batch(10)
    assert n == ["x", "y"] and b.n == 2 and b.seq.tolist() == [0, 1, 2, 3, 2, 2]
    p.write_text("")
    b, n = FastqStream(str(p)).next_batch(10)
    assert b.n == 0 and n == []
    p.write_text("@x\nACGT\n+\n")
    with pytest.raises(RuntimeError):
        FastqStream(str(p)).next_batch(10)


def test_interleave():
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_1.fq"), 10 ** 9)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "rep_reads_2.fq"), 10 ** 9)
    a, _ = FastqStream(os.path.join(GOLDEN, "rep_reads_1.fq")).next_batch(1000)
    b, _ = FastqStream(os.path.join(GOLDEN, "rep_reads_2.fq")).next_batch(1000)
    il = interleave(a, b)
    assert il.n == 2 * len(r1)
    want_seq = np.concatenate([x for p in zip(r1, r2) for x in p])
    want_q = np.concatenate([x for p in zip(q1, q2) for x in p])
    assert np.array_equal(il.seq, want_seq) and np.array_equal(il.qual, want_q)
    assert np.array_equal(il.lengths(), np.array([len(x) for p in zip(r1, r2) for x in p]))
    c, _ = FastqStream(os.path.join(GOLDEN, "rep_reads_2.fq")).next_batch(10)
    with pytest.raises(ValueError):
        interleave(a, c)


def test_align_files_fails_loudly_without_gpu(tmp_path, lambda_index):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bowtie2_b200 import Bt2GpuError
    out = tmp_path / "o.sam"
    with pytest.raises(Bt2GpuError):
        align_files(lambda_index, str(out), os.path.join(GOLDEN, "lambda_reads_1.fq"))
    assert not out.exists()


@pytest.mark.parametrize("paired", [False, True])
def test_exact_mode_whole_files_over_a_fake_device(tmp_path, lambda_index, paired):
    """align_files(exact=True): FASTQ in -> policy engine in waves -> SAM out, with the CPU oracle answering behind the GPU
    entry-point conventions (tests/fake_gpu.py): header, records and order identical to the reference program's file."""
    from fake_gpu import FakeGpu
    from oracle_lib import Oracle
    golden_path = os.path.join(GOLDEN, "lambda_P_sensitive.sam" if paired else "lambda_U_sensitive.sam")
    r1, r2 = os.path.join(GOLDEN, "lambda_reads_1.fq"), os.path.join(GOLDEN, "lambda_reads_2.fq")
    out = str(tmp_path / "o.sam")
    import io
    summ = io.StringIO()
    if paired:
        for src, dst in ((r1, "a.fq"), (r2, "b.fq")):
            with open(src) as f, open(tmp_path / dst, "w") as g:
                g.writelines(f.readlines()[:800])                 # the golden holds the first 200 pairs
        align_files(lambda_index, out, str(tmp_path / "a.fq"), str(tmp_path / "b.fq"), exact=True, batch_reads=150, summary=summ,
                    gpu=FakeGpu(Oracle(lambda_index)))
    else:
        with open(r1) as f, open(tmp_path / "a.fq", "w") as g:
            g.writelines(f.readlines()[:4 * 500])
        align_files(lambda_index, out, str(tmp_path / "a.fq"), exact=True, batch_reads=128, summary=summ, gpu=FakeGpu(Oracle(lambda_index)))
    want = [l for l in open(golden_path)]
    got = [l for l in open(out)]
    n = len(got)
    assert got == want[:n] and n == (402 if paired else 502)        # 2 header lines + the records
    if paired:
        # the summary text equals the reference's except for the documented concordant ">1" split (0 here either way)
        assert summ.getvalue() == open(os.path.join(GOLDEN, "lambda_P_sensitive.summary.txt")).read()


def test_exact_mode_k_hits_whole_files_over_a_fake_device(tmp_path, rep_index):
    """align_files(exact=True, policy_options={"k": 3}) on the repeat-rich fixture: every record of -k 3 (FLAG 256 secondaries
    included) and the alignment summary identical to the reference program's"""
    import io
    import subprocess
    from fake_gpu import FakeGpu
    from oracle_lib import Oracle, have_reference, ref_bin
    if not have_reference():
        pytest.skip("oracle/_ref not built")
    fq = os.path.join(GOLDEN, "rep_reads_1.fq")
    with open(fq) as f, open(tmp_path / "a.fq", "w") as g:
        g.writelines(f.readlines()[:4 * 300])
    p = subprocess.run([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-k", "3", "-x", rep_index, "-U", str(tmp_path / "a.fq")],
                       capture_output=True, text=True, check=True)
    want = [l for l in p.stdout.split("\n") if l and not l.startswith("@")]
    out, summ = str(tmp_path / "o.sam"), io.StringIO()
    align_files(rep_index, out, str(tmp_path / "a.fq"), exact=True, batch_reads=128, summary=summ, gpu=FakeGpu(Oracle(rep_index)),
                policy_options={"k": 3})
    got = [l.rstrip("\n") for l in open(out) if not l.startswith("@")]
    assert got == want, next((a, b) for a, b in zip(got, want) if a != b)
    assert sum(int(l.split("\t")[1]) & 256 != 0 for l in want) > 50
    assert summ.getvalue() == p.stderr


def test_exact_mode_paired_k_hits_whole_files_over_a_fake_device(tmp_path, rep_index):
    """align_files(exact=True, policy_options={"k": 3}) on PAIRS of the repeat-rich fixture (bt2g_policy_align_pairs_k): every record of
    -k 3 -- secondary pairs and the mates' secondary alignments included -- identical to the reference program's"""
    import io
    import subprocess
    from fake_gpu import FakeGpu
    from oracle_lib import Oracle, have_reference, ref_bin
    if not have_reference():
        pytest.skip("oracle/_ref not built")
    for m in (1, 2):
        with open(os.path.join(GOLDEN, f"rep_reads_{m}.fq")) as f, open(tmp_path / f"a{m}.fq", "w") as g:
            g.writelines(f.readlines()[:4 * 200])
    p = subprocess.run([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-k", "3", "-x", rep_index,
                        "-1", str(tmp_path / "a1.fq"), "-2", str(tmp_path / "a2.fq")], capture_output=True, text=True, check=True)
    want = [l for l in p.stdout.split("\n") if l and not l.startswith("@")]
    out, summ = str(tmp_path / "o.sam"), io.StringIO()
    align_files(rep_index, out, str(tmp_path / "a1.fq"), str(tmp_path / "a2.fq"), exact=True, batch_reads=128, summary=summ,
                gpu=FakeGpu(Oracle(rep_index)), policy_options={"k": 3})
    got = [l.rstrip("\n") for l in open(out) if not l.startswith("@")]
    assert len(got) == len(want) and got == want, next(((a, b) for a, b in zip(got, want) if a != b), (len(got), len(want)))
    assert sum(int(l.split("\t")[1]) & 256 != 0 for l in want) > 50


@pytest.mark.parametrize("flags,options", [(["--no-discordant"], {"discord": False}), (["--no-mixed"], {"mixed": False}),
                                           (["--no-discordant", "--no-mixed"], {"discord": False, "mixed": False})])
def test_exact_mode_no_discordant_no_mixed_whole_files(tmp_path, rep_index, flags, options):
    """align_files(exact=True, policy_options={"discord": False} / {"mixed": False}) on pairs of the repeat-rich fixture: records AND
    the alignment summary identical to the reference program's (--no-discordant has to reach the record formatter and the counts:
    a pair whose mates each aligned once is then two unpaired alignments, YT:Z:UP)"""
    import io
    import subprocess
    from fake_gpu import FakeGpu
    from oracle_lib import Oracle, have_reference, ref_bin
    if not have_reference():
        pytest.skip("oracle/_ref not built")
    for m in (1, 2):
        with open(os.path.join(GOLDEN, f"rep_reads_{m}.fq")) as f, open(tmp_path / f"a{m}.fq", "w") as g:
            g.writelines(f.readlines()[:4 * 300])
    p = subprocess.run([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", *flags, "-x", rep_index,
                        "-1", str(tmp_path / "a1.fq"), "-2", str(tmp_path / "a2.fq")], capture_output=True, text=True, check=True)
    want = [l for l in p.stdout.split("\n") if l and not l.startswith("@")]
    plain = subprocess.run([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", rep_index,
                            "-1", str(tmp_path / "a1.fq"), "-2", str(tmp_path / "a2.fq")], capture_output=True, text=True, check=True)
    assert plain.stdout != p.stdout                                  # (the options matter on this sample)
    out, summ = str(tmp_path / "o.sam"), io.StringIO()
    align_files(rep_index, out, str(tmp_path / "a1.fq"), str(tmp_path / "a2.fq"), exact=True, batch_reads=128, summary=summ,
                gpu=FakeGpu(Oracle(rep_index)), policy_options=options)
    got = [l.rstrip("\n") for l in open(out) if not l.startswith("@")]
    assert len(got) == len(want) and got == want, next(((a, b) for a, b in zip(got, want) if a != b), (len(got), len(want)))
    # the summary: every line but the documented split of the concordant pairs into "exactly 1" / ">1" (the
    # engines keep one pair per read); those two lines must add up to the same number
    def split(text):
        conc, rest = 0, []
        for l in text.split("\n"):
            if "aligned concordantly exactly 1 time" in l or "aligned concordantly >1 times" in l:
                conc += int(l.split()[0])
            else:
                rest.append(l)
        return conc, rest
    assert split(summ.getvalue()) == split(p.stderr), (summ.getvalue(), p.stderr)
