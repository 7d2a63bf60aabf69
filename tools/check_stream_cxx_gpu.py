"""For the next GPU session (`tools/gpu_session.sh <tag> streamcxx`): bt2g_stream_run over two DEVICE engines + a solo engine on the golden
lambda pairs (blocks mode and byte-stream mode); the SAM text must equal the reference program's golden file.  Written after the round's
GPU budget was spent: the same loop is pinned on the CPU (tests/test_stream.py); this is the device-side check that is still owed, kept
out of the pytest suite until it has run once."""
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import subprocess
    import tempfile
    from bowtie2_b200 import Bt2Gpu
    from bowtie2_b200.lib import XEngine, load_library, policy_params, stream_run
    golden_dir = os.path.join(ROOT, "tests", "golden")
    d = tempfile.mkdtemp()
    base = os.path.join(d, "lambda")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "bowtie2-build-s"), "-q", os.path.join(golden_dir, "lambda_virus.fa"), base], stdout=subprocess.DEVNULL)
    g = Bt2Gpu(0)
    g.load_index_files(base)
    golden = [l.rstrip("\n") for l in open(os.path.join(golden_dir, "lambda_P_sensitive.sam")) if not l.startswith("@")]
    n = len(golden) // 2
    t1 = b"".join(open(os.path.join(golden_dir, "lambda_reads_1.fq"), "rb").readlines()[:4 * n])
    t2 = b"".join(open(os.path.join(golden_dir, "lambda_reads_2.fq"), "rb").readlines()[:4 * n])
    l1, l2 = t1.split(b"\n"), t2.split(b"\n")
    cut = 700
    blocks = [(b"\n".join(l1[4 * a:4 * min(a + cut, n)]) + b"\n", b"\n".join(l2[4 * a:4 * min(a + cut, n)]) + b"\n") for a in range(0, n, cut)]
    lib = load_library()
    engines = [XEngine(g, policy_params("sensitive", paired=True), cut, 512) for _ in range(2)]
    solo = XEngine(g, policy_params("sensitive"), 64, 512)
    ok = True
    for what, kw in (("blocks", dict(blocks=blocks)), ("byte streams", dict(blocks=None, files=[io.BytesIO(t1), io.BytesIO(t2)], chunk_bytes=1 << 16))):
        chunks = []
        kw = dict(kw)
        written, rc, counts = stream_run(lib, engines, kw.pop("blocks"), chunks.append, ["gi|9626243|ref|NC_001416.1|"], paired=True, max_units=cut, max_len=512,
                                         max_ops=512 + 80, name_stride=64, want_counts=True, solo=solo, **kw)
        lines = b"".join(chunks).decode().rstrip("\n").split("\n")
        bad = [i for i in range(len(golden)) if i >= len(lines) or lines[i] != golden[i]]
        print(f"bt2g_stream_run over device engines, {what}: {written} records, rc {rc}, {len(golden) - len(bad)} of {len(golden)} identical to the reference program")
        ok = ok and not bad and rc == 0 and written == len(golden)
    for e in engines + [solo]:
        e.close()
    g.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
