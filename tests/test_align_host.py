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
