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


# Every instrumented binary of this file is compiled when the first test asks for one, all at the same time (g++ -fsanitize spends ~10 - 30 s
# on each: side by side they cost the suite one compile, not five); a test then waits for its own.
_BUILDS, _DONE = {}, {}


def _start_builds():
    if _BUILDS:
        return
    import tempfile
    from oracle_lib import ref_bin
    d = tempfile.mkdtemp(prefix="bt2g_sanitize_")
    csrc, san_dir, inc = os.path.join(ROOT, "bowtie2_b200", "csrc"), os.path.join(ROOT, "tests", "sanitize"), os.path.join(ROOT, "include")
    base = ["g++", "-O1", "-g", "-fno-omit-frame-pointer", "-I", inc]
    specs = {
        "fuzz_parse": base + ["-fsanitize=address,undefined", "-std=c++17", os.path.join(san_dir, "fuzz_parse.cpp"), os.path.join(csrc, "sam_host.cpp"), "-lpthread"],
        "fuzz_format": base + ["-fsanitize=address,undefined", "-std=c++17", os.path.join(san_dir, "fuzz_format.cpp"), os.path.join(csrc, "sam_host.cpp"), "-lpthread"],
        "run_engines": base + ["-fsanitize=address,undefined", "-std=c++20", "-ffp-contract=off", os.path.join(san_dir, "run_engines.cpp"),
                               os.path.join(csrc, "xengine_host.cpp"), os.path.join(csrc, "policy_engine.cpp"), ref_bin("libbt2oracle.so"),
                               "-Wl,-rpath," + os.path.dirname(ref_bin("libbt2oracle.so")), "-lpthread",
                               "-Wl,--unresolved-symbols=ignore-all"],   # (bt2g_policy_backend_gpu names the device entry points; nothing here calls it)
    }
    for san in ("address,undefined", "thread"):
        specs["run_stream:" + san] = base + ["-fsanitize=" + san, "-std=c++17", os.path.join(san_dir, "run_stream.cpp"), os.path.join(csrc, "stream_host.cpp"),
                                             os.path.join(csrc, "sam_host.cpp"), "-lpthread"]
    for k, (name, cmd) in enumerate(specs.items()):
        exe = os.path.join(d, f"bin{k}")
        _BUILDS[name] = (subprocess.Popen(cmd + ["-o", exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True), exe)


def _built(name):
    """path of the instrumented binary `name` (skips without a sanitizer runtime)"""
    _start_builds()
    if name not in _DONE:
        p, exe = _BUILDS[name]
        _, err = p.communicate()
        _DONE[name] = (p.returncode, err, exe)
    rc, err, exe = _DONE[name]
    if rc and ("asan" in err or "tsan" in err or "sanitize" in err):
        pytest.skip("no sanitizer runtime for g++ here")
    assert rc == 0, err[-3000:]
    return exe


def _build(tmp_path, name):
    return _built(name)


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


def _build_engines(tmp_path):
    return _built("run_engines")


def _engines_case(exe, seed, k, tmp_path, n_unpaired=120, n_pairs=60):
    """one tests/parity_fuzz.py case through the uninstrumented library (against the reference program) and through the instrumented build;
    False for the cases the state machine does not take (-k / -a)"""
    import ctypes as C
    import parity_fuzz
    from oracle_lib import Oracle, oracle_policy_table
    lib = load_library()
    c = parity_fuzz.draw_case(seed, k)
    if c["kw"].get("k") is not None or c["kw"].get("all_hits"):
        return False
    work = str(tmp_path / f"c{seed}_{k}")
    n, bad, first, st, desc = parity_fuzz.run_case(c, work, n_unpaired=n_unpaired, n_pairs=n_pairs)
    assert bad == 0, desc
    # the same inputs for the instrumented build: reads as the harness made them, parameters as lib.policy_params packs them
    local, paired, large = c["local"], c["paired"], c.get("large", False)
    files = [os.path.join(work, "r1.fq"), os.path.join(work, "r2.fq")] if paired else [os.path.join(work, "r.fq")]
    per = [read_fastq_codes(f, 10 ** 6) for f in files]
    if paired:
        N = [x for pr in zip(per[0][0], per[1][0]) for x in pr]; R = [x for pr in zip(per[0][1], per[1][1]) for x in pr]; Q = [x for pr in zip(per[0][2], per[1][2]) for x in pr]
    else:
        N, R, Q = per[0]
    batch = ReadBatch.from_list(R, Q)
    prm = policy_params(c["preset"], local=local, paired=paired, seed=c.get("run_seed", 0), **c["kw"])
    be, keep = oracle_policy_table(Oracle(os.path.join(work, "g")), local, 8 if large else 4, c["kw"].get("sc"))
    res, ops, pairs, _ = policy_align(lib, be, prm, batch, N, entry="bt2g_xengine_align_host", max_ops=4 * int(batch.lengths().max()) + 64)
    stride = 64
    names = np.zeros((batch.n, stride), dtype=np.uint8)
    for i, s in enumerate(N):
        b = s.encode()[:stride - 1]
        names[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    sc = c["kw"].get("sc")
    sci = np.array([sc.match_bonus, sc.mmp_max, sc.mmp_min, sc.n_pen, sc.rdgap_const, sc.rdgap_linear, sc.rfgap_const, sc.rfgap_linear] if sc else [-1] * 8, dtype=np.int32)
    scd = np.array([sc.n_ceil_over.C, sc.n_ceil_over.L] if sc is not None and sc.n_ceil_over is not None else [0.0, -1.0], dtype=np.float64)
    dump = os.path.join(work, "engines.bin")
    with open(dump, "wb") as f:
        f.write(np.array([batch.n, int(batch.off[-1]), stride, C.sizeof(prm), int(local), 8 if large else 4], dtype=np.uint64).tobytes())
        f.write(sci.tobytes()); f.write(scd.tobytes()); f.write(bytes(prm))
        for a in (batch.seq, batch.qual, batch.off, names, res):
            f.write(np.ascontiguousarray(a).tobytes())
    p = subprocess.run([exe, os.path.join(work, "g"), dump], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert p.returncode == 0, (desc, p.stdout[-800:], p.stderr[-4000:])
    return True


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.timeout(900)
def test_search_policy_engines_under_sanitizers(tmp_path):
    """csrc/xengine.cuh (the state machine the device runs) and csrc/policy_engine.cpp built with ASan + UBSan and driven over the C oracle's
    table on a few tests/parity_fuzz.py cases (local / end-to-end, paired / unpaired, .bt2l, odd reads, cheap-gap scoring, small seeds):
    no out-of-bounds access of the fixed-capacity unit state, same results as the uninstrumented library"""
    import parity_fuzz
    if not os.path.exists(parity_fuzz.REF):
        pytest.skip("oracle/_ref is not built")
    exe = _build_engines(tmp_path)
    ran = sum(_engines_case(exe, seed, k, tmp_path) for seed, k in ((114, 3), (114, 8), (101, 13), (31, 5), (31, 12), (64, 7), (43, 2), (44, 64)))
    assert ran >= 5


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.timeout(600)
@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_cxx_stream_under_sanitizers(tmp_path, san):
    """bt2g_stream_run (csrc/stream_host.cpp) with synthetic engines that finish out of order: the SAM text equals the blocks formatted one
    after the other, a failing engine / reader / writer ends the run with its code and without a hang; under ASan + UBSan, and under TSan
    (reader, E engine threads and the writer share the slot queues)."""
    exe = _built("run_stream:" + san)
    for seed in (1, 2):
        r = subprocess.run([exe, str(seed), "60"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "0 inconsistencies" in r.stdout, (r.stdout[-600:], r.stderr[-3000:])
