#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
Outputs (all small, committed):
    lambda_virus.fa              example reference genome (data fixture, copied verbatim)
    lambda_reads_{1,2}.fq        first 2000 read pairs of example/reads/reads_{1,2}.fq
    lambda_fm_golden.npz         Ebwt::countBt2SideEx / mapLF1 / ftabLoHi / getOffset /
                                 joinedToTextOff answers of the reference for seeded queries,
                                 plus exactSweep and searchAllSeeds results for the first 300 reads
    lambda_U_sensitive.sam       bowtie2-align-s --sensitive -U lambda_reads_1.fq --seed 0 (config 1)
    lambda_U_local.sam           --local --sensitive-local for the first 300 reads
    lambda_P_sensitive.sam       the same for the first 200 pairs of lambda_reads_{1,2}.fq (-1/-2)
The index itself is rebuilt at test time with oracle/_ref/bowtie2-build-s (default parameters);
the generating command is recorded in the npz.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Reference, ref_bin  # noqa: E402

REF = "/root/reference"
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def read_fastq(path, n):
    out = []
    with open(path) as f:
        for _ in range(n):
            name = f.readline().strip()
            seq = f.readline().strip()
            f.readline()
            q = f.readline().strip()
            if not q:
                break
            out.append((name, seq, q))
    return out


def main():
    shutil.copy(os.path.join(REF, "example/reference/lambda_virus.fa"), os.path.join(HERE, "lambda_virus.fa"))
    for m in (1, 2):
        recs = read_fastq(os.path.join(REF, f"example/reads/reads_{m}.fq"), 2000)
        with open(os.path.join(HERE, f"lambda_reads_{m}.fq"), "w") as f:
            for name, seq, q in recs:
                f.write(f"{name}\n{seq}\n+\n{q}\n")
    tmp = tempfile.mkdtemp()
    base = os.path.join(tmp, "lambda")
    cmd = [ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", os.path.join(HERE, "lambda_virus.fa"), base]
    subprocess.check_call(cmd)
    R = Reference(base)
    sc = R.scalars()
    rng = np.random.default_rng(7)
    n = sc["bwt_len"]
    g = {}
    for m, tag in ((False, "fw"), (True, "bw")):
        zo = R.scalars(m)["z_off"]
        rows = np.concatenate([rng.integers(0, n, 4000), [0, n - 1, zo, min(zo + 1, n - 1), max(zo, 1) - 1]]).astype(np.uint64)
        g[f"rank_rows_{tag}"] = rows
        g[f"rank4_{tag}"] = R.rank4(rows, m)
        chars = rng.integers(0, 4, len(rows)).astype(np.uint8)
        g[f"lf1_chars_{tag}"] = chars
        g[f"lf1_{tag}"] = R.maplf1(rows, chars, m)
        idx = rng.integers(0, sc["ftab_len"] - 1, 3000).astype(np.uint64)
        g[f"ftab_idx_{tag}"] = idx
        g[f"ftab_{tag}"] = R.ftab_lohi(idx, m)
    rows = rng.integers(0, n, 3000).astype(np.uint64)
    g["off_rows"] = rows
    g["offsets"] = R.get_offset(rows)
    jt = []
    for off in g["offsets"][:1000]:
        for qlen in (1, 20, 250):
            if off + qlen > sc["len"]:
                continue
            ok, ti, to, tl, st = R.joined_to_text(qlen, int(off), 1)
            jt.append((int(off), qlen, ok, ti, to, tl, st))
    g["joined"] = np.array(jt, dtype=np.uint64)
    recs = read_fastq(os.path.join(HERE, "lambda_reads_1.fq"), 300)
    sweep, seeds = [], []
    MS = 32
    for name, seq, q in recs:
        codes = np.array([CODE.get(c, 4) for c in seq], dtype=np.uint8)
        nelt, mine, tb = R.exact_sweep(codes)
        sweep.append([nelt] + mine + tb)
        ln = len(codes)
        interval = max(1, int(1 + 1.15 * np.sqrt(ln)))   # --sensitive: S,1,1.15
        nn, out = R.seed_search(codes, 22, interval, 0, MS)
        seeds.append(out)
    g["sweep"] = np.array(sweep, dtype=np.uint64)
    g["seeds"] = np.stack(seeds)
    g["seed_params"] = np.array([22, MS], dtype=np.int64)
    g["build_cmd"] = np.array(" ".join(["bowtie2-build-s", "--seed", "0", "--quiet", "lambda_virus.fa", "lambda"]))
    np.savez_compressed(os.path.join(HERE, "lambda_fm_golden.npz"), **g)
    sam = os.path.join(HERE, "lambda_U_sensitive.sam")
    subprocess.check_call([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base,
                           "-U", os.path.join(HERE, "lambda_reads_1.fq"), "-S", sam], stderr=subprocess.DEVNULL)
    # drop the @PG line (contains paths)
    lines = [l for l in open(sam) if not l.startswith("@PG")]
    open(sam, "w").writelines(lines)
    # local mode: the first 300 reads
    sam = os.path.join(HERE, "lambda_U_local.sam")
    subprocess.check_call([ref_bin("bowtie2-align-s"), "--local", "--sensitive-local", "--seed", "0", "-p", "1", "-x", base, "-u", "300",
                           "-U", os.path.join(HERE, "lambda_reads_1.fq"), "-S", sam], stderr=subprocess.DEVNULL)
    lines = [l for l in open(sam) if not l.startswith("@PG")]
    open(sam, "w").writelines(lines)
    # paired: the first 200 pairs (config 3 flags: FR, -I 0 -X 500)
    sam = os.path.join(HERE, "lambda_P_sensitive.sam")
    subprocess.check_call([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-u", "200",
                           "-1", os.path.join(HERE, "lambda_reads_1.fq"), "-2", os.path.join(HERE, "lambda_reads_2.fq"),
                           "-S", sam], stderr=subprocess.DEVNULL)
    lines = [l for l in open(sam) if not l.startswith("@PG")]
    open(sam, "w").writelines(lines)
    shutil.rmtree(tmp)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
