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
    rep_genome.fa, rep_reads_{1,2}.fq, rep_{U,P}_sensitive.sam
                                 synthetic repeat-rich set (see repeat_fixture) and the reference's output for it
    *.summary.txt                the alignment summary the reference printed on stderr for each of the three runs
                                 (AlnSink::printAlSumm, aln_sink.cpp:349-528)
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


def repeat_fixture():
    """A 24 kbp two-contig genome with six repeat families, 400 read pairs of 90 bp with substitutions and indels, some
    of them too far apart, on the same strand, or with one unalignable mate: exercises XS:i / low MAPQ, discordant
    and unpaired-mate records and the ">1 times" lines of the summary, which the lambda set never reaches."""
    rng=np.random.default_rng(20260923)
    G=rng.integers(0,4,24000).astype(np.uint8)
    # repeat families: 6 families x 3-4 copies of 400 bp, 0-3 substitutions per copy
    for fam in range(6):
        src=rng.integers(0,len(G)-400)
        seg=G[src:src+400].copy()
        for c in range(int(rng.integers(2,4))):
            dst=rng.integers(0,len(G)-400)
            s=seg.copy()
            for _ in range(int(rng.integers(0,4))):
                p=rng.integers(0,400); s[p]=(s[p]+1+rng.integers(0,3))%4
            G[dst:dst+400]=s
    # two contigs
    ctg=[G[:14000],G[14000:]]
    with open(f'{HERE}/rep_genome.fa','w') as f:
        for i,c in enumerate(ctg):
            f.write(f'>ctg{i+1} synthetic\n')
            s=''.join('ACGT'[x] for x in c)
            for k in range(0,len(s),70): f.write(s[k:k+70]+'\n')
    comp=np.array([3,2,1,0],dtype=np.uint8)
    def mutate(s):
        s=list(s)
        out=[]
        for ch in s:
            u=rng.random()
            if u<0.01: out.append((ch+1+rng.integers(0,3))%4)
            elif u<0.012: continue
            elif u<0.014: out.append(ch); out.append(rng.integers(0,4))
            else: out.append(ch)
        return np.array(out,dtype=np.uint8)
    def qual(n):
        q=np.clip(40-np.arange(n)*20//n+rng.integers(-3,4,n),2,41)
        return ''.join(chr(33+x) for x in q)
    f1=open(f'{HERE}/rep_reads_1.fq','w'); f2=open(f'{HERE}/rep_reads_2.fq','w')
    NP=400
    for i in range(NP):
        ci=int(rng.integers(0,2)); c=ctg[ci]
        if rng.random()<0.03:
            a=rng.integers(0,4,90).astype(np.uint8); b=rng.integers(0,4,90).astype(np.uint8)
        else:
            kind=rng.random()
            frag=int(np.clip(rng.normal(300,40),150,480))
            if kind<0.06: frag=int(rng.integers(650,1200))          # too long for -X 500: discordant
            st=int(rng.integers(0,len(c)-frag))
            fr=c[st:st+frag]
            a=fr[:90]; b=comp[fr[-90:][::-1]]
            if 0.06<=kind<0.10: b=fr[-90:]                           # same strand: not concordant under --fr
            if 0.10<=kind<0.14: b=rng.integers(0,4,90).astype(np.uint8)   # mate 2 unalignable
            if 0.14<=kind<0.16: a=rng.integers(0,4,90).astype(np.uint8)   # mate 1 unalignable
            if rng.random()<0.5: a,b=b,a
            a=mutate(a); b=mutate(b)
        for f,s in ((f1,a),(f2,b)):
            f.write(f'@p{i}\n'+''.join('ACGT'[x] for x in s)+'\n+\n'+qual(len(s))+'\n')
    f1.close(); f2.close()
    tmp = tempfile.mkdtemp()
    base = tmp + "/rep"
    subprocess.check_call([ref_bin("bowtie2-build-s"), "--seed", "0", "--quiet", f"{HERE}/rep_genome.fa", base])
    for tag, args in (("U", ["-U", f"{HERE}/rep_reads_1.fq"]),
                      ("P", ["-1", f"{HERE}/rep_reads_1.fq", "-2", f"{HERE}/rep_reads_2.fq"])):
        sam = f"{HERE}/rep_{tag}_sensitive.sam"
        subprocess.check_call([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base] + args + ["-S", sam],
                              stderr=open(sam[:-4] + ".summary.txt", "w"))
        lines = [l for l in open(sam) if not l.startswith("@PG")]
        open(sam, "w").writelines(lines)
    shutil.rmtree(tmp)


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
                           "-U", os.path.join(HERE, "lambda_reads_1.fq"), "-S", sam], stderr=open(sam[:-4] + ".summary.txt", "w"))
    # drop the @PG line (contains paths)
    lines = [l for l in open(sam) if not l.startswith("@PG")]
    open(sam, "w").writelines(lines)
    # local mode: the first 300 reads
    sam = os.path.join(HERE, "lambda_U_local.sam")
    subprocess.check_call([ref_bin("bowtie2-align-s"), "--local", "--sensitive-local", "--seed", "0", "-p", "1", "-x", base, "-u", "300",
                           "-U", os.path.join(HERE, "lambda_reads_1.fq"), "-S", sam], stderr=open(sam[:-4] + ".summary.txt", "w"))
    lines = [l for l in open(sam) if not l.startswith("@PG")]
    open(sam, "w").writelines(lines)
    # paired: the first 200 pairs (config 3 flags: FR, -I 0 -X 500)
    sam = os.path.join(HERE, "lambda_P_sensitive.sam")
    subprocess.check_call([ref_bin("bowtie2-align-s"), "--sensitive", "--seed", "0", "-p", "1", "-x", base, "-u", "200",
                           "-1", os.path.join(HERE, "lambda_reads_1.fq"), "-2", os.path.join(HERE, "lambda_reads_2.fq"),
                           "-S", sam], stderr=open(sam[:-4] + ".summary.txt", "w"))
    lines = [l for l in open(sam) if not l.startswith("@PG")]
    open(sam, "w").writelines(lines)
    shutil.rmtree(tmp)
    repeat_fixture()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
