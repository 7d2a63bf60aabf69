"""Host stages of the text path on synthetic 2x150 bp pairs, without a GPU: FASTQ text -> interleaved batch (two bt2g_fastq_parse_mt
calls + the numpy interleave, fresh buffers: the first host path; against bt2g_fastq_parse_pairs_mt into reused buffers) and result
arrays -> SAM text (bytes from fresh buffers against a view of a reused buffer).  Fabricated gapless results with one mismatch: the
formatter's common case.  Prints one JSON line; `python tools/host_text_bench.py [threads] [pairs]`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bowtie2_b200.lib import PAIR_RESULT, READ_RESULT, HostBuffers, fastq_parse, fastq_parse_pairs, load_library, sam_format  # noqa: E402
from bowtie2_b200.stream import interleave_uniform  # noqa: E402


def fastq_text(rng, n, L, tag):
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, L))]
    q = rng.integers(35, 74, (n, L)).astype(np.uint8)
    hdr = np.frombuffer(b"".join(("@read%09d/%s\n" % (i, tag)).encode() for i in range(n)), dtype=np.uint8)
    W = len(hdr) // n
    buf = np.empty((n, W + L + 3 + L + 1), dtype=np.uint8)
    buf[:, :W] = hdr.reshape(n, W)
    buf[:, W:W + L] = seq; buf[:, W + L] = 10; buf[:, W + L + 1] = ord("+"); buf[:, W + L + 2] = 10
    buf[:, W + L + 3:W + 2 * L + 3] = q; buf[:, -1] = 10
    return buf.tobytes()


def best(f, reps=5):
    b, r = 1e9, None
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); b = min(b, time.perf_counter() - t0)
    return b, r


def main():
    TH = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    L = 150
    lib = load_library()
    rng = np.random.default_rng(1)
    T1, T2 = fastq_text(rng, N, L, "1"), fastq_text(rng, N, L, "2")

    def first_path():
        b1, n1, _ = fastq_parse(lib, T1, name_stride=32, threads=TH); b2, n2, _ = fastq_parse(lib, T2, name_stride=32, threads=TH)
        return interleave_uniform(b1, b2, n1, n2)
    slot = HostBuffers()
    t_old, (bo, no) = best(first_path)
    t_new, (bn, nn, _, _) = best(lambda: fastq_parse_pairs(lib, T1, T2, name_stride=32, threads=TH, out=slot))
    assert np.array_equal(bo.seq, bn.seq) and np.array_equal(bo.off, bn.off) and np.array_equal(no.rows, nn.rows) and np.array_equal(bo.qual, bn.qual)
    R = np.zeros(2 * N, dtype=READ_RESULT); R["score2"] = -(1 << 31)
    R["found"] = 1; R["nops"] = L; R["fw"] = rng.integers(0, 2, 2 * N); R["refoff"] = rng.integers(0, 40000, 2 * N); R["mapq"] = 42; R["score"] = -5
    O = np.zeros((2 * N, L + 64), dtype=np.uint8); O[:, 7] = 1 | (2 << 2)
    P = np.zeros(N, dtype=PAIR_RESULT); P["pair_type"] = 1
    out = HostBuffers()
    f_old, so = best(lambda: sam_format(lib, bn, R, O, ["chr1"], read_names=nn, pairs=P, threads=TH, as_bytes=True))
    f_new, sn = best(lambda: sam_format(lib, bn, R, O, ["chr1"], read_names=nn, pairs=P, threads=TH, as_bytes="view", out=out))
    assert bytes(sn) == so
    per = 1e6 / N
    print(json.dumps({"what": "host stages of the text path, seconds per 1 M 2x150 bp pairs (best of 5), no GPU involved", "threads": TH, "pairs": N,
                      "fastq_MB": (len(T1) + len(T2)) / 1e6, "sam_MB": len(so) / 1e6,
                      "parse": {"two_parses_plus_numpy_interleave_fresh_buffers": t_old * per, "bt2g_fastq_parse_pairs_mt_reused_buffers": t_new * per},
                      "format": {"bytes_from_fresh_buffers": f_old * per, "view_of_a_reused_buffer": f_new * per},
                      "note": "the first column still runs today's parser / formatter code underneath: the original first path was slower than it",
                      "host": os.uname().nodename, "cpus": os.cpu_count()}))


if __name__ == "__main__":
    main()
