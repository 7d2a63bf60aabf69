"""Deterministic synthetic genomes and reads.

Genome: i.i.d. uniform ACGT contigs plus a "repeat" track (random segments copied many times)
so multi-hit seeds and the -M logic are exercised, and one N gap per contig so the rstarts
fragment table is exercised.  Reads: uniform positions, both strands, substitutions + short
indels, a fraction of unalignable random reads; Phred qualities from a linear-decay profile.
"""
from __future__ import annotations

import numpy as np

DNA = np.frombuffer(b"ACGTN", dtype=np.uint8)
COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def make_genome(n_contigs: int, contig_len: int, seed: int = 20260922, repeat_frac: float = 0.01,
                repeat_len: int = 500, repeat_copies: int = 20, n_gap: int = 50):
    """Returns list of uint8 code arrays (0..3, 4 = N)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    contigs = [rng.integers(0, 4, size=contig_len, dtype=np.uint8) for _ in range(n_contigs)]
    total = n_contigs * contig_len
    n_fam = max(1, int(total * repeat_frac / (repeat_len * repeat_copies))) if repeat_frac > 0 else 0
    for _ in range(n_fam):
        src_c = int(rng.integers(0, n_contigs))
        if contig_len <= repeat_len + 1:
            break
        src_p = int(rng.integers(0, contig_len - repeat_len))
        seg = contigs[src_c][src_p:src_p + repeat_len].copy()
        for _ in range(repeat_copies):
            c = int(rng.integers(0, n_contigs))
            p = int(rng.integers(0, contig_len - repeat_len))
            contigs[c][p:p + repeat_len] = seg
    if n_gap > 0:
        for c in contigs:
            if len(c) > 4 * n_gap:
                p = len(c) // 2
                c[p:p + n_gap] = 4
    return contigs


def write_fasta(path: str, contigs, prefix: str = "chr"):
    with open(path, "wb") as f:
        for i, c in enumerate(contigs):
            f.write(f">{prefix}{i + 1}\n".encode())
            s = DNA[c].tobytes()
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + b"\n")


def revcomp(codes: np.ndarray) -> np.ndarray:
    return COMP[codes[::-1]]


def make_reads(contigs, n_reads: int, read_len: int, seed: int = 1, sub_rate: float = 0.005,
               indel_rate: float = 0.0005, random_frac: float = 0.01):
    """Returns (reads list of code arrays, quals list of Phred+33 arrays, truth list)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    reads, quals, truth = [], [], []
    base_q = np.linspace(40, 20, read_len)
    for i in range(n_reads):
        if rng.random() < random_frac:
            r = rng.integers(0, 4, size=read_len, dtype=np.uint8)
            truth.append((-1, -1, 0))
        else:
            c = int(rng.integers(0, len(contigs)))
            g = contigs[c]
            span = read_len + 8
            p = int(rng.integers(0, max(1, len(g) - span)))
            src = g[p:p + span].copy()
            out = []
            k = 0
            while len(out) < read_len and k < len(src):
                u = rng.random()
                if u < indel_rate / 2:            # deletion from read
                    k += 1 + int(rng.geometric(0.5)) - 1
                    continue
                if u < indel_rate:                # insertion into read
                    for _ in range(int(rng.geometric(0.5))):
                        out.append(int(rng.integers(0, 4)))
                    continue
                b = int(src[k])
                if b < 4 and rng.random() < sub_rate:
                    b = (b + 1 + int(rng.integers(0, 3))) % 4
                out.append(b)
                k += 1
            r = np.array(out[:read_len], dtype=np.uint8)
            if len(r) < read_len:
                r = np.concatenate([r, rng.integers(0, 4, size=read_len - len(r), dtype=np.uint8)])
            fw = rng.random() < 0.5
            if not fw:
                r = revcomp(r)
            truth.append((c, p, 1 if fw else -1))
        q = np.clip(base_q + rng.normal(0, 3, read_len), 2, 41).astype(np.uint8) + 33
        reads.append(r)
        quals.append(q)
    return reads, quals, truth


def write_fastq(path: str, reads, quals, prefix: str = "r"):
    with open(path, "wb") as f:
        for i, (r, q) in enumerate(zip(reads, quals)):
            f.write(f"@{prefix}{i}\n".encode())
            f.write(DNA[r].tobytes() + b"\n+\n" + q.tobytes() + b"\n")


def make_pairs(contigs, n_pairs: int, read_len: int, seed: int = 1, sub_rate: float = 0.005, indel_rate: float = 0.0005,
               ins_mean: float = 350.0, ins_sd: float = 30.0, hard_frac: float = 0.0, hard_period: int = 14):
    """FR pairs: fragment length ~ N(ins_mean, ins_sd) clipped to [read_len + 20, 500];
    mate 1 is the fragment's left end on a random strand of the fragment, mate 2 the reverse complement of its
    right end.  A `hard_frac` share of pairs gets a substitution every `hard_period` bases in one mate, which
    leaves no exact seed there so that only the mate-finding DP can place it.
    Returns (reads interleaved m1, m2, ..., quals, truth [(contig, fragment start, fragment length, flipped)])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    reads, quals, truth = [], [], []
    base_q = np.linspace(40, 20, read_len)

    def mutate(src):
        out, k = [], 0
        while len(out) < read_len and k < len(src):
            u = rng.random()
            if u < indel_rate / 2:
                k += 1
                continue
            if u < indel_rate:
                out.append(int(rng.integers(0, 4)))
                continue
            b = int(src[k])
            if b < 4 and rng.random() < sub_rate:
                b = (b + 1 + int(rng.integers(0, 3))) % 4
            out.append(b)
            k += 1
        r = np.array(out[:read_len], dtype=np.uint8)
        if len(r) < read_len:
            r = np.concatenate([r, rng.integers(0, 4, size=read_len - len(r), dtype=np.uint8)])
        return r

    for i in range(n_pairs):
        c = int(rng.integers(0, len(contigs)))
        g = contigs[c]
        frag = int(np.clip(rng.normal(ins_mean, ins_sd), read_len + 20, 500))
        p = int(rng.integers(0, max(1, len(g) - frag - 16)))
        left = mutate(g[p:p + read_len + 8])
        right = mutate(revcomp(g[p + frag - read_len - 8:p + frag]))     # reads the fragment's right end inwards
        flipped = rng.random() < 0.5                                      # fragment from the reverse strand
        m1, m2 = (right, left) if flipped else (left, right)
        if rng.random() < hard_frac:
            tgt = m2 if rng.random() < 0.5 else m1
            for k in range(int(rng.integers(0, hard_period)), read_len, hard_period):
                if tgt[k] < 4:
                    tgt[k] = (tgt[k] + 1 + int(rng.integers(0, 3))) % 4
        for m in (m1, m2):
            reads.append(m)
            quals.append(np.clip(base_q + rng.normal(0, 3, read_len), 2, 41).astype(np.uint8) + 33)
        truth.append((c, p, frag, flipped))
    return reads, quals, truth
