"""Host ceiling of the whole text loop, without a GPU: FASTQ text -> SAM text around STAND-IN engines that fabricate their results at
once (gapless alignments with one mismatch, concordant pairs: the formatter's common case), so that only the host stages are timed:
reader + parse || hand-over || format + counts + writer.  bt2g_stream_run (csrc/stream_host.cpp, C++ threads; blocks mode and
byte-stream mode) beside stream.TextAligner (Python threads over the same C calls).  Prints one JSON line;
`python tools/host_stream_bench.py [host threads] [pairs per block] [blocks]`."""
import ctypes as C
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from bowtie2_b200 import lib as L  # noqa: E402
from bowtie2_b200.stream import TextAligner  # noqa: E402
from host_text_bench import fastq_text  # noqa: E402


def main():
    TH = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    E, RL, NS = 2, 150, 32
    lib = L.load_library()
    rng = np.random.default_rng(1)
    t1, t2 = fastq_text(rng, B, RL, "1"), fastq_text(rng, B, RL, "2")
    max_ops = RL + 64
    # fabricated results of one block, written by the stand-in engines into whatever buffers they are given
    R = np.zeros(2 * B, dtype=L.READ_RESULT); R["score2"] = -(1 << 31)
    R["found"] = 1; R["nops"] = RL; R["fw"] = rng.integers(0, 2, 2 * B); R["refoff"] = rng.integers(0, 40000, 2 * B); R["mapq"] = 42; R["score"] = -5
    O = np.zeros((2 * B, max_ops), dtype=np.uint8); O[:, 7] = 1 | (2 << 2)
    P = np.zeros(B, dtype=L.PAIR_RESULT); P["pair_type"] = 1
    pthr = max(1, TH * 3 // 8)
    fthr = max(1, TH - pthr - E)

    # ---- C++ loop
    def cb(_eng, reads, _names, _stride, res, ops, mo, pairs, _stats):
        n = int(reads.contents.n_reads)
        C.memmove(res, R.ctypes.data, n * L.READ_RESULT.itemsize)
        C.memmove(ops, O.ctypes.data, n * mo)
        C.memmove(pairs, P.ctypes.data, (n // 2) * L.PAIR_RESULT.itemsize)
        return 0
    fn = L._STREAM_ALIGN(cb)
    lib.bt2g_stream_run.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.POINTER(L._StreamParams), C.POINTER(L._SamOpts), C.POINTER(L._StreamIO),
                                    C.c_void_p, C.POINTER(C.c_uint64), C.c_char_p, C.c_uint32]
    rn = (C.c_char_p * 1)(b"chr1")
    opt = L._SamOpts(rn, 1, None, fthr, 0, 0.0, 0.0, 0, 0, None)
    handles = (C.c_void_p * E)(*[None] * E)
    sam_bytes = [0]

    def write(_u, _p, n):
        sam_bytes[0] += n
        return 0

    def run_cxx(streamed):
        left = [K]
        files = [io.BytesIO(t1 * K), io.BytesIO(t2 * K)] if streamed else None

        def next_block(_u, a, la, b, lb):
            if left[0] == 0:
                return 0
            left[0] -= 1
            a[0], la[0] = C.cast(C.c_char_p(t1), C.c_void_p).value, len(t1)
            b[0], lb[0] = C.cast(C.c_char_p(t2), C.c_void_p).value, len(t2)
            return 1

        def read(_u, mate, dst, cap):
            return files[mate].readinto((C.c_char * cap).from_address(dst)) or 0
        sp = L._StreamParams(1, pthr, fthr, 2, B, RL, max_ops, NS, 0, 64 << 20, None, 0)
        sio = L._StreamIO(None, L._STREAM_NEXT(next_block), L._STREAM_WRITE(write), L._STREAM_READ(read) if streamed else L._STREAM_READ())
        if streamed:
            sio.next_block = L._STREAM_NEXT()
        counts = np.zeros(1, dtype=L.ALIGN_COUNTS)
        n, err = C.c_uint64(0), C.create_string_buffer(256)
        sam_bytes[0] = 0
        t0 = time.perf_counter()
        rc = lib.bt2g_stream_run(C.cast(fn, C.c_void_p), handles, E, C.byref(sp), C.byref(opt), C.byref(sio), counts.ctypes.data, C.byref(n), err, 256)
        dt = time.perf_counter() - t0
        assert rc == 0 and n.value == 2 * B * K and int(counts["nconcord_uni1"][0]) == B * K, (rc, err.value, n.value)
        return dt

    # ---- Python-thread loop over the same C calls
    class Fake:
        def align(self, batch, names, out=None):
            return R[:batch.n], O[:batch.n], P[:batch.n // 2], None

    def run_py():
        ta = TextAligner([Fake() for _ in range(E)], ["chr1"], True, parse_threads=pthr, format_threads=fthr, name_stride=NS)
        sam_bytes[0] = 0

        def sink(v):
            sam_bytes[0] += len(v)
        ta.run(((t1, t2) for _ in range(2)), sink)          # warm the buffer sets
        sam_bytes[0] = 0
        t0 = time.perf_counter()
        recs = ta.run(((t1, t2) for _ in range(K)), sink)
        dt = time.perf_counter() - t0
        assert recs == 2 * B * K
        return dt

    run_cxx(False)                                           # warm-up (page faults of the first buffers)
    d_blocks = min(run_cxx(False) for _ in range(3)); sam = sam_bytes[0]
    d_stream = min(run_cxx(True) for _ in range(3))
    d_py = min(run_py() for _ in range(3))
    pairs = B * K
    print(json.dumps({"what": "host ceiling of the whole text loop (FASTQ text -> SAM text, 2x150 bp pairs) around stand-in engines that fabricate their "
                              "results at once; Mpairs/s, best of 3; no GPU involved",
                      "host_threads": TH, "parse_threads": pthr, "format_threads": fthr, "engines": E, "pairs": pairs, "fastq_MB": (len(t1) + len(t2)) * K / 1e6,
                      "sam_MB": sam / 1e6,
                      "bt2g_stream_run_blocks": pairs / d_blocks / 1e6, "bt2g_stream_run_byte_streams": pairs / d_stream / 1e6,
                      "stream_TextAligner_python_threads": pairs / d_py / 1e6, "host": os.uname().nodename, "cpus": os.cpu_count()}))


if __name__ == "__main__":
    main()
