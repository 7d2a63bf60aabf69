"""bench.py's host-side helpers (no GPU): the FASTQ sample the reference arm and the parity gate read, the read names the engine
hashes into the per-read RNG seed (they must spell the FASTQ's), the reference's thread-count candidates, the SAM reader."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_fastq_sample_and_device_names_agree(tmp_path):
    rng = np.random.default_rng(3)
    n, L = 37, 25
    reads = rng.integers(0, 5, (n, L), dtype=np.uint8)
    quals = (rng.integers(2, 41, (n, L)) + 33).astype(np.uint8)
    p = str(tmp_path / "s.fq")
    bench.write_fastq(p, reads, quals, first_id=1234)
    text = open(p, "rb").read()
    assert text == bench.fastq_text(reads, quals, first_id=1234)
    lines = text.decode().split("\n")
    assert len(lines) == 4 * n + 1 and lines[-1] == ""
    rows = bench.device_name_rows(torch, torch.device("cpu"), 1234, n, 1).numpy()
    for i in range(n):
        name = bytes(rows[i]).split(b"\0")[0].decode()
        assert lines[4 * i] == "@" + name and name == "r%09d" % (1234 + i)
        assert lines[4 * i + 1] == "".join("ACGTN"[c] for c in reads[i]) and lines[4 * i + 2] == "+"
        assert lines[4 * i + 3].encode() == bytes(quals[i])
    # pairs: both mates of a pair carry the pair's name
    rows2 = bench.device_name_rows(torch, torch.device("cpu"), 0, 3, 2).numpy()
    assert [bytes(r).split(b"\0")[0] for r in rows2] == [b"r000000000", b"r000000000", b"r000000001", b"r000000001", b"r000000002", b"r000000002"]
    # the library's parser reads the sample back
    from bowtie2_b200.lib import fastq_parse, load_library
    b, names, used = fastq_parse(load_library(), text, name_stride=16)
    assert b.n == n and used == len(text) and np.array_equal(b.seq.reshape(n, L), reads) and np.array_equal(b.qual.reshape(n, L), quals)
    assert list(names)[0] == "r%09d" % 1234


def test_reference_thread_candidates():
    assert bench.reference_thread_candidates(128, 16.0) == [32, 16]          # the container's quota and twice that
    assert bench.reference_thread_candidates(128, 96.0) == [128, 96]
    assert bench.reference_thread_candidates(8, None) == [8, 4]              # no quota: every hardware thread and half
    assert bench.reference_thread_candidates(4, 16.0) == [4]


def test_sam_records_reader(tmp_path):
    p = str(tmp_path / "x.sam")
    open(p, "w").write("@HD\tVN:1.0\n@SQ\tSN:chr1\tLN:10\n@SQ\tSN:chr2\tLN:20\n@PG\tID:x\nr1\t4\t*\nr2\t0\tchr2\t5\n")
    recs, names = bench.sam_records(p)
    assert recs == ["r1\t4\t*", "r2\t0\tchr2\t5"] and names == ["chr1", "chr2"]
