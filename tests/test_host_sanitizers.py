"""The host C++ of the library's text boundary (csrc/sam_host.cpp: FASTQ parsers, SAM formatter) under AddressSanitizer +
UndefinedBehaviorSanitizer: tests/sanitize/fuzz_parse.cpp and fuzz_format.cpp feed it random / damaged inputs with exact-size buffers
and check that the multi-threaded entry points agree with the serial ones."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from bowtie2_b200.lib import PAIR_RESULT, READ_RESULT, ReadBatch, load_library, policy_align, policy_params
from conftest import GOLDEN, read_fastq_codes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name):
    exe = str(tmp_path / name)
    cmd = ["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "sanitize", name + ".cpp"), os.path.join(ROOT, "bowtie2_b200", "csrc", "sam_host.cpp"), "-lpthread"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode and ("asan" in p.stderr or "sanitize" in p.stderr):
        pytest.skip("no sanitizer runtime for g++ here")
    assert p.returncode == 0, p.stderr[-2000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.timeout(600)
def test_fastq_parsers_under_sanitizers(tmp_path):
    exe = _build(tmp_path, "fuzz_parse")
    for seed in (1, 5):
        p = subprocess.run([exe, str(seed), "120"], capture_output=True, text=True)
        assert p.returncode == 0 and "0 inconsistencies" in p.stdout, (p.stdout[-600:], p.stderr[-3000:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.timeout(600)
def test_sam_formatter_under_sanitizers(tmp_path, lambda_index):
    from oracle_lib import Oracle, oracle_policy_table
    lib = load_library()
    n = 250
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_2.fq"), n)
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    R, Q, N = il(r1, r2), il(q1, q2), il(n1, n2)
    be, keep = oracle_policy_table(Oracle(lambda_index), False, 4)
    batch = ReadBatch.from_list(R, Q)
    res, ops, pairs, _ = policy_align(lib, be, policy_params("sensitive", paired=True), batch, N, entry="bt2g_xengine_align_host")
    stride = 64
    names = np.zeros((2 * n, stride), dtype=np.uint8)
    for i, s in enumerate(N):
        b = s.encode()[:stride - 1]
        names[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    dump = str(tmp_path / "data.bin")
    with open(dump, "wb") as f:
        f.write(np.array([2 * n, ops.shape[1], stride, int(batch.off[-1]), READ_RESULT.itemsize, PAIR_RESULT.itemsize], dtype=np.uint64).tobytes())
        for a in (batch.seq, batch.qual, batch.off, res, ops, pairs, names):
            f.write(np.ascontiguousarray(a).tobytes())
    exe = _build(tmp_path, "fuzz_format")
    for seed in (1, 2):
        p = subprocess.run([exe, str(seed), dump], capture_output=True, text=True)
        assert p.returncode == 0 and "0 inconsistencies" in p.stdout, (p.stdout[-600:], p.stderr[-3000:])
