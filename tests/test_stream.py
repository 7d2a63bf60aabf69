"""bowtie2_b200/stream.py: FASTQ text in -> SAM text out with parse / align / format overlapped on host threads.  The engines
here are CPU stand-ins (the device engine's state machine driven over the oracle-backed table, bt2g_xengine_align_host); the SAM
must equal the reference program's golden file whatever the batch cut and the number of engines."""
import os

import pytest

from bowtie2_b200.lib import load_library, policy_align, policy_params
from bowtie2_b200.stream import TextAligner
from conftest import GOLDEN
from test_policy_engine_cpp import _table


import threading

_ORACLE_LOCK = threading.Lock()          # the oracle-backed table is Python callbacks over one C oracle: one caller at a time


class _HostEngine:
    def __init__(self, be, params):
        self.be, self.params, self.lib = be, params, load_library()

    def align(self, batch, names):
        with _ORACLE_LOCK:
            res, ops, pairs, st = policy_align(self.lib, self.be, self.params, batch, names, entry="bt2g_xengine_align_host")
        return res, ops, pairs, st


def _records(path, n):
    """the first n FASTQ records of a file as bytes"""
    out, k = [], 0
    with open(path, "rb") as f:
        for line in f:
            out.append(line)
            k += 1
            if k == 4 * n:
                break
    return out


@pytest.mark.parametrize("n_engines,cut", [(1, 100), (2, 37), (3, 64)])
def test_fastq_text_to_sam_text_paired(n_engines, cut, lambda_index):
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_P_sensitive.sam")) if not l.startswith("@")]
    n = 200
    l1, l2 = _records(os.path.join(GOLDEN, "lambda_reads_1.fq"), n), _records(os.path.join(GOLDEN, "lambda_reads_2.fq"), n)
    items = [(b"".join(l1[4 * a:4 * min(a + cut, n)]), b"".join(l2[4 * a:4 * min(a + cut, n)])) for a in range(0, n, cut)]
    be, keep, fake = _table(lambda_index)
    engines = [_HostEngine(be, policy_params("sensitive", paired=True)) for _ in range(n_engines)]
    ta = TextAligner(engines, ["gi|9626243|ref|NC_001416.1|"], paired=True, parse_threads=2, format_threads=2, name_stride=64)
    chunks = []
    written = ta.run(iter(items), lambda v: chunks.append(bytes(v)))   # (the view is valid only inside the sink)
    lines = b"".join(chunks).decode().rstrip("\n").split("\n")
    assert written == 2 * n and lines == golden[:2 * n]


def test_fastq_text_to_sam_text_unpaired(lambda_index):
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_U_sensitive.sam")) if not l.startswith("@")]
    n = 300
    l1 = _records(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    items = [(b"".join(l1[4 * a:4 * min(a + 128, n)]), None) for a in range(0, n, 128)]
    be, keep, fake = _table(lambda_index)
    ta = TextAligner([_HostEngine(be, policy_params("sensitive")) for _ in range(2)], ["gi|9626243|ref|NC_001416.1|"], paired=False,
                     parse_threads=1, format_threads=1, name_stride=64)
    chunks = []
    ta.run(iter(items), lambda v: chunks.append(bytes(v)))
    assert b"".join(chunks).decode().rstrip("\n").split("\n") == golden[:n]


@pytest.mark.parametrize("paired,gz", [(True, False), (True, True), (False, False)])
def test_whole_files_through_the_stream(tmp_path, lambda_index, paired, gz):
    """stream.align_files_stream: FASTQ FILES (plain / .gz, mate files read in step, ragged batch cuts) -> SAM file + alignment summary
    around two engines; header, records and summary equal the reference program's golden outputs"""
    import gzip
    import io
    import shutil
    from bowtie2_b200.stream import align_files_stream
    fix = "lambda_P_sensitive" if paired else "lambda_U_sensitive"
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, fix + ".sam")) if not l.startswith("@")]
    files = []
    for m in ((1, 2) if paired else (1,)):
        src = os.path.join(GOLDEN, f"lambda_reads_{m}.fq")
        n_rec = len(golden) // (2 if paired else 1)
        lines = open(src, "rb").readlines()[:4 * n_rec]
        dst = str(tmp_path / f"r{m}.fq") + (".gz" if gz else "")
        with (gzip.open(dst, "wb") if gz else open(dst, "wb")) as f:
            f.writelines(lines)
        files.append(dst)
    be, keep, fake = _table(lambda_index)
    made = []

    def make_engine(prm, max_units, max_len):
        assert max_units == 333 and max_len == 600
        made.append(_HostEngine(be, prm))
        return made[-1]
    out, summ = str(tmp_path / "o.sam"), io.StringIO()
    counts = align_files_stream(lambda_index, out, files[0], files[1] if paired else None, preset="sensitive", engines=2, batch_units=333,
                                max_read_len=600, threads=3, summary=summ, gpu=object(), make_engine=make_engine)
    got = [l.rstrip("\n") for l in open(out)]
    assert [l for l in got if l.startswith("@")][:2] == ["@HD\tVN:1.5\tSO:unsorted\tGO:query", "@SQ\tSN:gi|9626243|ref|NC_001416.1|\tLN:48502"]
    assert [l for l in got if not l.startswith("@")] == golden and len(made) == 2
    assert summ.getvalue() == open(os.path.join(GOLDEN, fix + ".summary.txt")).read()
    assert int(counts["nread"][0]) == len(golden) // (2 if paired else 1)


def test_stream_files_errors(tmp_path, lambda_index):
    """a mate file that ends early and a read name longer than the name rows are errors, not silent cuts"""
    from bowtie2_b200.stream import align_files_stream
    l1 = open(os.path.join(GOLDEN, "lambda_reads_1.fq"), "rb").readlines()[:400]
    l2 = open(os.path.join(GOLDEN, "lambda_reads_2.fq"), "rb").readlines()[:360]
    (tmp_path / "a1.fq").write_bytes(b"".join(l1)); (tmp_path / "a2.fq").write_bytes(b"".join(l2))
    be, keep, fake = _table(lambda_index)
    mk = lambda prm, n, l: _HostEngine(be, prm)
    with pytest.raises(RuntimeError, match="fewer reads in file specified with -2"):
        align_files_stream(lambda_index, str(tmp_path / "o.sam"), str(tmp_path / "a1.fq"), str(tmp_path / "a2.fq"), batch_units=64, gpu=object(), make_engine=mk)
    (tmp_path / "b.fq").write_bytes(b"@" + b"x" * 200 + b"\nACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIII\n")
    with pytest.raises(ValueError, match="name_stride"):
        align_files_stream(lambda_index, str(tmp_path / "o.sam"), str(tmp_path / "b.fq"), batch_units=64, gpu=object(), make_engine=mk)


# ---- the same loop in C++ (csrc/stream_host.cpp: bt2g_stream_run) -----------------------------------------------------------
def _host_align(be_by_engine, params):
    lib = load_library()

    def align(j, batch, names):
        with _ORACLE_LOCK:
            res, ops, pairs, _ = policy_align(lib, be_by_engine[j], params, batch, names, entry="bt2g_xengine_align_host")
        return res, ops, pairs
    return align


@pytest.mark.parametrize("n_engines,cut", [(1, 100), (2, 37), (3, 64)])
def test_cxx_stream_paired(n_engines, cut, lambda_index):
    """bt2g_stream_run: reader callback + parse, one thread per engine, ordered formatter + writer callback; the SAM text and the
    alignment counts equal the reference program's whatever the block cut and the number of engines"""
    from bowtie2_b200.lib import align_summary, stream_run
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_P_sensitive.sam")) if not l.startswith("@")]
    n = 200
    l1, l2 = _records(os.path.join(GOLDEN, "lambda_reads_1.fq"), n), _records(os.path.join(GOLDEN, "lambda_reads_2.fq"), n)
    blocks = [(b"".join(l1[4 * a:4 * min(a + cut, n)]), b"".join(l2[4 * a:4 * min(a + cut, n)])) for a in range(0, n, cut)]
    be, keep, fake = _table(lambda_index)
    lib = load_library()
    chunks = []
    written, rc, counts = stream_run(lib, list(range(n_engines)), blocks, chunks.append, ["gi|9626243|ref|NC_001416.1|"], paired=True, max_units=cut,
                                     max_len=1024, max_ops=1024 + 64, name_stride=64, align=_host_align([be] * n_engines, policy_params("sensitive", paired=True)),
                                     want_counts=True)
    lines = b"".join(chunks).decode().rstrip("\n").split("\n")
    assert written == 2 * n and rc == 0 and lines == golden[:2 * n]
    # the counts of the same records through the Python-side helper
    from bowtie2_b200.lib import ReadBatch, align_counts_add
    from conftest import read_fastq_codes
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_2.fq"), n)
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    res, ops, pairs, _ = policy_align(lib, be, policy_params("sensitive", paired=True), ReadBatch.from_list(il(r1, r2), il(q1, q2)), il(n1, n2),
                                      entry="bt2g_xengine_align_host")
    assert align_summary(lib, counts) == align_summary(lib, align_counts_add(lib, None, res, pairs))


def test_cxx_stream_unpaired_and_errors(lambda_index):
    from bowtie2_b200.lib import stream_run
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_U_sensitive.sam")) if not l.startswith("@")]
    n = 300
    l1 = _records(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    blocks = [(b"".join(l1[4 * a:4 * min(a + 128, n)]), None) for a in range(0, n, 128)]
    be, keep, fake = _table(lambda_index)
    lib = load_library()
    kw = dict(paired=False, max_units=128, max_len=1024, max_ops=1088, name_stride=64, parse_threads=1, format_threads=1,
              align=_host_align([be, be], policy_params("sensitive")))
    chunks = []
    written, rc, _ = stream_run(lib, [0, 1], blocks, chunks.append, ["gi|9626243|ref|NC_001416.1|"], **kw)
    assert written == n and rc == 0 and b"".join(chunks).decode().rstrip("\n").split("\n") == golden[:n]
    # no input: no records, no error
    assert stream_run(lib, [0, 1], [], chunks.append, ["gi|9626243|ref|NC_001416.1|"], **kw)[:2] == (0, 0)
    # a block that ends inside a record, a block larger than the engines' capacity, a read longer than max_len, a failing sink
    with pytest.raises(RuntimeError, match="whole records"):
        stream_run(lib, [0], [(b"".join(l1[:7]), None)], chunks.append, ["x"], **kw)
    with pytest.raises(RuntimeError, match="whole records"):
        stream_run(lib, [0], [(b"".join(l1[:4 * 129]), None)], chunks.append, ["x"], **kw)
    with pytest.raises(RuntimeError, match="max_len"):
        stream_run(lib, [0], blocks, chunks.append, ["x"], **{**kw, "max_len": 20})

    def bad_sink(_):
        raise OSError("disk full")
    with pytest.raises(OSError):
        stream_run(lib, [0, 1], blocks, bad_sink, ["gi|9626243|ref|NC_001416.1|"], **kw)
    # a pair with an empty mate 2 is refused (the reference aligns its mate 1 as an unpaired read: stream.TextAligner does that)
    pe = dict(kw, paired=True, align=_host_align([be], policy_params("sensitive", paired=True)))
    with pytest.raises(RuntimeError, match="empty mate 2"):
        stream_run(lib, [0], [(b"@a\nACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIII\n", b"@a\n\n+\n\n")], chunks.append, ["x"], **pe)


@pytest.mark.parametrize("paired,gz,chunk", [(True, False, 5000), (True, True, 700), (False, False, 0)])
def test_cxx_stream_whole_files(tmp_path, lambda_index, paired, gz, chunk):
    """bt2g_stream_run with the read callback: the mate FILES (plain / .gz) as byte streams, the library's reader cuts the blocks (records
    cut by the chunk end, chunks smaller than a record, mate files read in step by record); records and alignment summary equal the
    reference program's golden outputs"""
    import gzip
    from bowtie2_b200.lib import align_summary, stream_run
    fix = "lambda_P_sensitive" if paired else "lambda_U_sensitive"
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, fix + ".sam")) if not l.startswith("@")]
    files = []
    for m in ((1, 2) if paired else (1,)):
        n_rec = len(golden) // (2 if paired else 1)
        lines = open(os.path.join(GOLDEN, f"lambda_reads_{m}.fq"), "rb").readlines()[:4 * n_rec]
        lines[-1] = lines[-1].rstrip(b"\n")                      # (no final newline)
        dst = str(tmp_path / f"r{m}.fq") + (".gz" if gz else "")
        with (gzip.open(dst, "wb") if gz else open(dst, "wb")) as f:
            f.writelines(lines)
        files.append(gzip.open(dst, "rb") if gz else open(dst, "rb"))
    be, keep, fake = _table(lambda_index)
    lib = load_library()
    chunks = []
    written, rc, counts = stream_run(lib, [0, 1], None, chunks.append, ["gi|9626243|ref|NC_001416.1|"], paired=paired, max_units=333, max_len=1024,
                                     max_ops=1088, name_stride=128, align=_host_align([be, be], policy_params("sensitive", paired=paired)), want_counts=True,
                                     files=files, chunk_bytes=chunk)
    for f in files:
        f.close()
    assert rc == 0 and written == len(golden) and b"".join(chunks).decode().rstrip("\n").split("\n") == golden
    assert align_summary(lib, counts) == open(os.path.join(GOLDEN, fix + ".summary.txt")).read()


def test_cxx_stream_file_errors(tmp_path, lambda_index):
    """a mate file that ends early, input that ends inside a record and a read name longer than the name rows are errors, not silent cuts"""
    import io
    from bowtie2_b200.lib import stream_run
    l1 = open(os.path.join(GOLDEN, "lambda_reads_1.fq"), "rb").readlines()[:400]
    l2 = open(os.path.join(GOLDEN, "lambda_reads_2.fq"), "rb").readlines()[:360]
    be, keep, fake = _table(lambda_index)
    lib = load_library()
    kw = dict(max_units=64, max_len=1024, max_ops=1088, name_stride=64, chunk_bytes=3000)
    pe = dict(kw, paired=True, align=_host_align([be], policy_params("sensitive", paired=True)))
    se = dict(kw, paired=False, align=_host_align([be], policy_params("sensitive")))
    with pytest.raises(RuntimeError, match="fewer reads in file specified with -2"):
        stream_run(lib, [0], None, lambda b: None, ["x"], files=[io.BytesIO(b"".join(l1)), io.BytesIO(b"".join(l2))], **pe)
    with pytest.raises(RuntimeError, match="truncated FASTQ record"):
        stream_run(lib, [0], None, lambda b: None, ["x"], files=[io.BytesIO(b"".join(l1[:42]))], **se)
    with pytest.raises(RuntimeError, match="name_stride"):
        stream_run(lib, [0], None, lambda b: None, ["x"], files=[io.BytesIO(b"@" + b"x" * 200 + b"\nACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIII\n")], **se)
    # empty input
    assert stream_run(lib, [0], None, lambda b: None, ["x"], files=[io.BytesIO(b"")], **se)[:2] == (0, 0)
    assert stream_run(lib, [0], None, lambda b: None, ["x"], files=[io.BytesIO(b"\n"), io.BytesIO(b"")], **pe)[:2] == (0, 0)


def test_cxx_stream_pairs_with_an_empty_mate_2(lambda_index):
    """a pair whose mate 2 is empty is an unpaired read for the reference (bt2_search.cpp:3326): with a solo engine bt2g_stream_run writes
    its mate 1 as one unpaired record in place; same text and same counts as stream.TextAligner (whose handling of these pairs the file-level
    fuzz compares with the reference program), and the ordinary pairs around them equal the golden records"""
    from bowtie2_b200.lib import align_counts_add, align_summary, stream_run
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_P_sensitive.sam")) if not l.startswith("@")]
    n = 120
    l1, l2 = _records(os.path.join(GOLDEN, "lambda_reads_1.fq"), n), _records(os.path.join(GOLDEN, "lambda_reads_2.fq"), n)
    solos = {0, 17, 18, 63, 119}                                 # first, adjacent, at a block end, last
    for i in solos:
        l2[4 * i + 1], l2[4 * i + 3] = b"\n", b"\n"
    cut = 64
    blocks = [(b"".join(l1[4 * a:4 * min(a + cut, n)]), b"".join(l2[4 * a:4 * min(a + cut, n)])) for a in range(0, n, cut)]
    be, keep, fake = _table(lambda_index)
    lib = load_library()
    pp, up = policy_params("sensitive", paired=True), policy_params("sensitive")

    def align(j, batch, names):                                  # engines 0, 1: paired; 2: the solo engine
        with _ORACLE_LOCK:
            res, ops, pairs, _ = policy_align(lib, be, up if j == 2 else pp, batch, names, entry="bt2g_xengine_align_host")
        return res, ops, pairs
    chunks = []
    written, rc, counts = stream_run(lib, [0, 1], blocks, chunks.append, ["gi|9626243|ref|NC_001416.1|"], paired=True, max_units=cut, max_len=1024,
                                     max_ops=1088, name_stride=64, align=align, want_counts=True, solo=True, solo_max_units=2)
    got = b"".join(chunks).decode().rstrip("\n").split("\n")
    assert rc == 0 and written == 2 * n - len(solos) == len(got)
    # the Python-thread twin on the same blocks
    ta = TextAligner([_HostEngine(be, pp), _HostEngine(be, pp)], ["gi|9626243|ref|NC_001416.1|"], paired=True, parse_threads=2, format_threads=2, name_stride=64,
                     make_solo_engine=lambda: _HostEngine(be, up))
    want_chunks, cnt = [], [None]

    def on_batch(res, pairs):
        cnt[0] = align_counts_add(lib, cnt[0], res, pairs)
    ta.run(iter(blocks), lambda v: want_chunks.append(bytes(v)), on_batch=on_batch)
    want = b"".join(want_chunks).decode().rstrip("\n").split("\n")
    assert got == want
    assert align_summary(lib, counts) == align_summary(lib, cnt[0])
    # the ordinary pairs are the golden records; a solo leaves one record, flagged as an unpaired read
    gi = 0
    for i in range(n):
        if i in solos:
            rec = got[gi].split("\t")
            assert int(rec[1]) & 1 == 0 and "YT:Z:UU" in rec
            gi += 1
        else:
            assert got[gi:gi + 2] == golden[2 * i:2 * i + 2]
            gi += 2
    # without a solo engine such a pair is refused
    with pytest.raises(RuntimeError, match="empty mate 2"):
        stream_run(lib, [0], blocks, chunks.append, ["x"], paired=True, max_units=cut, max_len=1024, max_ops=1088, name_stride=64, align=align)
