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
