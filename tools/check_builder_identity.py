#!/usr/bin/env python
"""tools/check_builder_identity.py [--mbp 120] [--out profiles/r02_builder_identity.json]: the six index files written by the
bench's torch index builder (bowtie2_b200/index_build.py) against the files of the reference's own `bowtie2-build-s --seed 0` on
the same synthetic genome (repeat families, N gaps): sha256 of each file from both builders.  Needs a GPU (the builder) and
oracle/_ref (the reference builder); ~4 minutes at 120 Mbp on 16 host threads."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=120.0)
    ap.add_argument("--contigs", type=int, default=6)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_builder_identity.json"))
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from bowtie2_b200.index_build import build_index
    dev = torch.device("cuda", 0)
    clen = int(a.mbp * 1e6 / a.contigs)
    contigs = bench.make_genome_gpu(torch, dev, a.contigs, clen, repeat_fams=200)
    names = [f"chr{k + 1}" for k in range(a.contigs)]
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(d, "g.fa")
    dna = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(fa, "wb") as f:
        for nm, c in zip(names, contigs):
            f.write(b">" + nm.encode() + b"\n")
            s = dna[c.cpu().numpy()]
            w = 60
            body = np.full((len(s) + w - 1) // w * (w + 1), ord("\n"), dtype=np.uint8)
            idx = np.arange(len(s))
            body[idx + idx // w] = s
            f.write(body[:len(s) + (len(s) + w - 1) // w].tobytes())
    t0 = time.time()
    built = build_index(contigs, names=names, mirror_offs=True)     # (the bench skips the mirror index's SA sample: the aligner never reads it)
    torch.cuda.synchronize()
    ours = os.path.join(d, "ours")
    built.write_files(ours)
    t_ours = time.time() - t0
    t0 = time.time()
    ref = os.path.join(d, "ref")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "bowtie2-build-s"), "--seed", "0", "--quiet", "--threads", str(a.threads), fa, ref])
    t_ref = time.time() - t0
    files, same = {}, True
    for ext in ("1.bt2", "2.bt2", "3.bt2", "4.bt2", "rev.1.bt2", "rev.2.bt2"):
        ho, hr = sha(f"{ours}.{ext}"), sha(f"{ref}.{ext}")
        files[ext] = {"bytes": os.path.getsize(f"{ours}.{ext}"), "sha256_torch_builder": ho, "sha256_bowtie2_build_s": hr, "identical": ho == hr}
        same = same and ho == hr
    out = {"genome_mbp": a.contigs * clen / 1e6, "contigs": a.contigs, "identical": same, "files": files,
           "seconds_torch_builder_gpu": t_ours, "seconds_bowtie2_build_s": t_ref, "bowtie2_build_threads": a.threads}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "files"}))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
