#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json metric: Mreads/s, 2x150 bp, 3 Gbp index).

Default workload = BASELINE.json configs[2], the configuration the metric is quoted on: synthetic 3 Gbp genome
(.bt2 index built on the GPU by bowtie2_b200.index_build, byte-identical layout to bowtie2-build-s), 2x150 bp FR
pairs (fragment ~ N(350,30)), --end-to-end --very-sensitive.  A "read" in Mreads/s is one PAIR, as in the
reference's own summary ("N reads; of these: N were paired").  A "step" is one pass of the hot path over one batch
of `--batch` pairs: exactSweep -> multiseed search -> offset resolve -> extension DP + backtrace for both mates,
then mate framing -> mate-finding DP -> pair selection.  Batches are taken round-robin from a resident set.
`--workload se100` runs configs[1] (10 M x 100 bp unpaired, --sensitive) instead.

    python bench.py --gpus N --steps K --warmup W          # our arm (torchrun for N > 1)
    python bench.py --impl reference ...                    # the reference CPU bowtie2 on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the field definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_CONTIGS, CONTIG_LEN = 24, 125_000_000
WORKLOADS = {
    # name: paired, read length, preset, default number of units (pairs / reads) resident in HBM
    "pe150": dict(paired=True, read_len=150, preset="very-sensitive", units=10_000_000,
                  label="BASELINE.json configs[2]: 2x150 bp paired, --end-to-end --very-sensitive"),
    "se100": dict(paired=False, read_len=100, preset="sensitive", units=10_000_000,
                  label="BASELINE.json configs[1]: 1x100 bp unpaired, --end-to-end --sensitive"),
}
WORKDIR = os.environ.get("BT2G_BENCH_DIR", "/dev/shm/bt2g_bench")


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic data on the GPU
# ------------------------------------------------------------------------------------------------
def make_genome_gpu(torch, dev, n_contigs, contig_len, seed=20260922, repeat_fams=50, repeat_len=5000, repeat_copies=120,
                    n_gap=10_000):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    contigs = [torch.randint(0, 4, (contig_len,), dtype=torch.uint8, device=dev, generator=g) for _ in range(n_contigs)]
    rng = np.random.default_rng(seed)
    if contig_len > 4 * repeat_len:
        fams = max(1, int(repeat_fams * (n_contigs * contig_len) / 3e9))
        for _ in range(fams):
            sc, sp = int(rng.integers(0, n_contigs)), int(rng.integers(0, contig_len - repeat_len))
            seg = contigs[sc][sp:sp + repeat_len].clone()
            for _ in range(repeat_copies):
                c, p = int(rng.integers(0, n_contigs)), int(rng.integers(0, contig_len - repeat_len))
                contigs[c][p:p + repeat_len] = seg
    if contig_len > 8 * n_gap:
        for c in contigs:
            p = contig_len // 2
            c[p:p + n_gap] = 4
    return contigs


def make_reads_gpu(torch, dev, contigs, n_reads, read_len, seed=1, sub_rate=0.005, indel_frac=0.05, random_frac=0.01):
    """uint8 [n, L] codes and Phred+33 qualities on the device; reads are drawn from either strand."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    nc, clen = len(contigs), contigs[0].numel()
    genome = torch.cat(contigs)
    span = read_len + 4
    ci = torch.randint(0, nc, (n_reads,), device=dev, generator=g)
    pos = torch.randint(0, clen - span, (n_reads,), device=dev, generator=g)
    start = ci * clen + pos
    ar = torch.arange(read_len, device=dev)
    # one short indel in a fraction of reads (deletion from / insertion into the read)
    u = torch.rand(n_reads, device=dev, generator=g)
    dlen = torch.randint(1, 4, (n_reads,), device=dev, generator=g)
    ipos = torch.randint(10, read_len - 10, (n_reads,), device=dev, generator=g)
    is_del = u < indel_frac / 2
    is_ins = (u >= indel_frac / 2) & (u < indel_frac)
    shift = torch.zeros(n_reads, read_len, dtype=torch.int64, device=dev)
    after = ar[None, :] >= ipos[:, None]
    shift += (after & is_del[:, None]) * dlen[:, None]
    ins_amt = torch.clamp(ar[None, :] - ipos[:, None] + 1, min=0)
    ins_amt = torch.minimum(ins_amt, dlen[:, None])
    shift -= is_ins[:, None] * ins_amt
    idx = start[:, None] + ar[None, :] + shift
    reads = genome[idx]
    del idx, shift
    in_ins = is_ins[:, None] & after & (ar[None, :] < (ipos + dlen)[:, None])
    rnd_base = torch.randint(0, 4, (n_reads, read_len), dtype=torch.uint8, device=dev, generator=g)
    reads = torch.where(in_ins, rnd_base, reads)
    sub = (torch.rand(n_reads, read_len, device=dev, generator=g) < sub_rate) & (reads < 4)
    reads = torch.where(sub, (reads + 1 + rnd_base % 3) % 4, reads)
    randr = torch.rand(n_reads, device=dev, generator=g) < random_frac
    reads = torch.where(randr[:, None], rnd_base, reads)
    rc = torch.rand(n_reads, device=dev, generator=g) < 0.5
    comp = torch.tensor([3, 2, 1, 0, 4], dtype=torch.uint8, device=dev)
    reads = torch.where(rc[:, None], comp[reads.flip(1).long()], reads)
    q = torch.linspace(40, 20, read_len, device=dev)[None, :] + 3.0 * torch.randn(n_reads, read_len, device=dev, generator=g)
    quals = (torch.clamp(q, 2, 41).to(torch.uint8) + 33)
    del genome
    return reads.contiguous(), quals.contiguous()


def _with_errors(torch, dev, g, genome, start, read_len, sub_rate, indel_frac):
    """read_len bases starting at joined offset `start`, with substitutions and (in a fraction of reads) one short indel"""
    n = start.numel()
    ar = torch.arange(read_len, device=dev)
    u = torch.rand(n, device=dev, generator=g)
    dlen = torch.randint(1, 4, (n,), device=dev, generator=g)
    ipos = torch.randint(10, read_len - 10, (n,), device=dev, generator=g)
    is_del = u < indel_frac / 2
    is_ins = (u >= indel_frac / 2) & (u < indel_frac)
    after = ar[None, :] >= ipos[:, None]
    shift = (after & is_del[:, None]) * dlen[:, None]
    ins_amt = torch.minimum(torch.clamp(ar[None, :] - ipos[:, None] + 1, min=0), dlen[:, None])
    shift = shift - is_ins[:, None] * ins_amt
    reads = genome[start[:, None] + ar[None, :] + shift]
    del shift
    in_ins = is_ins[:, None] & after & (ar[None, :] < (ipos + dlen)[:, None])
    rnd_base = torch.randint(0, 4, (n, read_len), dtype=torch.uint8, device=dev, generator=g)
    reads = torch.where(in_ins, rnd_base, reads)
    sub = (torch.rand(n, read_len, device=dev, generator=g) < sub_rate) & (reads < 4)
    reads = torch.where(sub, (reads + 1 + rnd_base % 3) % 4, reads)
    return reads, rnd_base


def make_pairs_gpu(torch, dev, contigs, n_pairs, read_len, seed=1, sub_rate=0.005, indel_frac=0.05, random_frac=0.01,
                   ins_mean=350.0, ins_sd=30.0, chunk=1_000_000):
    """FR pairs: uint8 [2n, L] codes / qualities, mate 1 at even rows, mate 2 at odd rows."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    nc, clen = len(contigs), contigs[0].numel()
    genome = torch.cat(contigs)
    comp = torch.tensor([3, 2, 1, 0, 4], dtype=torch.uint8, device=dev)
    reads = torch.empty(2 * n_pairs, read_len, dtype=torch.uint8, device=dev)
    quals = torch.empty(2 * n_pairs, read_len, dtype=torch.uint8, device=dev)
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        frag = torch.clamp(ins_mean + ins_sd * torch.randn(n, device=dev, generator=g), read_len + 20, 500).long()
        ci = torch.randint(0, nc, (n,), device=dev, generator=g)
        pos = (torch.rand(n, device=dev, generator=g) * (clen - 520)).long()
        start = ci * clen + pos
        left, rnd_l = _with_errors(torch, dev, g, genome, start, read_len, sub_rate, indel_frac)
        right, rnd_r = _with_errors(torch, dev, g, genome, start + frag - read_len, read_len, sub_rate, indel_frac)
        right = comp[right.flip(1).long()]                      # mate from the fragment's right end reads inwards
        randp = torch.rand(n, device=dev, generator=g) < random_frac
        left = torch.where(randp[:, None], rnd_l, left)
        right = torch.where(randp[:, None], rnd_r, right)
        flip = torch.rand(n, device=dev, generator=g) < 0.5      # fragment taken from the reverse strand
        m1 = torch.where(flip[:, None], right, left)
        m2 = torch.where(flip[:, None], left, right)
        reads[2 * c0:2 * (c0 + n):2] = m1
        reads[2 * c0 + 1:2 * (c0 + n):2] = m2
        q = torch.linspace(40, 20, read_len, device=dev)[None, :] + 3.0 * torch.randn(2 * n, read_len, device=dev, generator=g)
        quals[2 * c0:2 * (c0 + n)] = torch.clamp(q, 2, 41).to(torch.uint8) + 33
        del left, right, rnd_l, rnd_r, m1, m2, q
    del genome
    return reads, quals


def write_fastq(path, reads_np, quals_np, first_id=0):
    """fixed-width FASTQ records written as one uint8 matrix (fast)."""
    n, L = reads_np.shape
    dna = np.frombuffer(b"ACGTN", dtype=np.uint8)
    idw = 9
    rec = 2 + idw + 1 + L + 3 + L + 1
    out = np.empty((n, rec), dtype=np.uint8)
    out[:, 0] = ord("@"); out[:, 1] = ord("r")
    ids = np.arange(first_id, first_id + n)
    for k in range(idw):
        out[:, 2 + idw - 1 - k] = (ids // 10 ** k) % 10 + ord("0")
    o = 2 + idw
    out[:, o] = ord("\n"); o += 1
    out[:, o:o + L] = dna[reads_np]; o += L
    out[:, o] = ord("\n"); out[:, o + 1] = ord("+"); out[:, o + 2] = ord("\n"); o += 3
    out[:, o:o + L] = quals_np; o += L
    out[:, o] = ord("\n")
    out.tofile(path)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on the host cores
# ------------------------------------------------------------------------------------------------
def ref_binary():
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    v256 = os.path.join(ROOT, "oracle", "_ref", "bowtie2-align-s-v256")
    sse = os.path.join(ROOT, "oracle", "_ref", "bowtie2-align-s")
    if " avx2 " in flags and " bmi2 " in flags and " fma " in flags and os.path.exists(v256):
        return v256, "bowtie2-align-s-v256 (AVX2)"
    return sse, "bowtie2-align-s (SSE2)"


def _ref_cmd(exe, preset, threads, index_base, fq):
    inp = ["-1", fq[0], "-2", fq[1]] if isinstance(fq, (tuple, list)) else ["-U", fq]
    return [exe, *preset, "--seed", "0", "-p", str(threads), "-x", index_base, *inp, "-S", "/dev/null"]


def time_reference(index_base, fq_small, fq_big, n_small, n_big, threads, preset):
    """reads (pairs)/s of the reference on the host cores, with index-load time removed by differencing
    two sample sizes (same command otherwise).  fq_* is a path (unpaired) or a (mate1, mate2) tuple."""
    exe, label = ref_binary()
    if not os.path.exists(exe):
        return None

    def run(fq):
        t0 = time.time()
        subprocess.check_call(_ref_cmd(exe, preset, threads, index_base, fq), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.time() - t0

    t_small = run(fq_small)
    t_big = run(fq_big)
    dt = max(t_big - t_small, 1e-6)
    return {"reads_per_s": (n_big - n_small) / dt, "t_small": t_small, "t_big": t_big, "binary": label, "threads": threads}


def best_thread_count(index_base, fq_small, n_small, cores, preset):
    """The reference does not always scale to every hardware thread (shared input/output locks);
    give it the thread count at which it is fastest on this box."""
    exe, _ = ref_binary()
    best, best_t = cores, None
    p = cores
    cands = []
    while p >= 8:
        cands.append(p)
        p //= 2
    for p in cands or [cores]:
        t0 = time.time()
        subprocess.check_call(_ref_cmd(exe, preset, p, index_base, fq_small), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.time() - t0
        log(f"reference -p {p}: {dt:.2f}s for {n_small} reads")
        if best_t is None or dt < best_t:
            best, best_t = p, dt
    return best


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pe150", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1_000_000, help="reads (pairs for a paired workload) per step")
    ap.add_argument("--genome-mbp", type=float, default=GENOME_CONTIGS * CONTIG_LEN / 1e6,
                    help="debug only: smaller genome (any value other than the default is NOT the BASELINE config)")
    ap.add_argument("--reads", type=int, default=0, help="reads (pairs) resident in HBM (0 = the workload's default)")
    ap.add_argument("--seed-table", type=int, default=16,
                    help="k of the extended seed table derived from the index at load time (0 = off; results are identical)")
    ap.add_argument("--dense-sa", type=int, default=0,
                    help="rate of the denser SA sample derived from the index at load time (0 = full suffix array, -1 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads (pairs) in the CPU baseline sample (0 = auto)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    paired, READ_LEN = wl["paired"], wl["read_len"]
    if args.reads <= 0:
        args.reads = wl["units"]
    mates = 2 if paired else 1
    ref_preset = ("--end-to-end", "--" + wl["preset"])

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1 and args.impl == "ours"
    if distributed:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    from bowtie2_b200 import Bt2Gpu
    from bowtie2_b200.index_build import build_index
    from bowtie2_b200.lib import Pipeline, READ_RESULT, PAIR_RESULT, _Reads

    full = abs(args.genome_mbp - GENOME_CONTIGS * CONTIG_LEN / 1e6) < 1e-6 and args.reads >= wl["units"]
    contig_len = int(args.genome_mbp * 1e6 / GENOME_CONTIGS)
    unit = "pairs" if paired else "reads"
    workload = (f"{wl['label']}: synthetic {GENOME_CONTIGS * contig_len / 1e9:.2f} Gbp genome .bt2 index, "
                f"{args.reads / 1e6:g}M {'2x' if paired else '1x'}{READ_LEN} bp {unit} resident in HBM")
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:                                   # container CPU quota, if any (explains where the reference stops scaling)
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        cpu_quota = None if q == "max" else float(q) / float(per)
    except Exception:
        cpu_quota = None

    # ---- setup (untimed): genome, index (built once on rank 0, NCCL-broadcast to the others), reads
    t0 = time.time()
    contigs = make_genome_gpu(torch, dev, GENOME_CONTIGS, contig_len)
    need_files = args.impl == "reference" or (rank == 0 and not args.no_cpu_baseline)
    built = None
    if rank == 0 or not distributed:
        built = build_index(contigs)
        torch.cuda.synchronize()
        log(f"rank {rank}: index built in {time.time() - t0:.1f}s (len={built.len})")
    gpu = Bt2Gpu(local_rank)
    bcast_s = 0.0
    if distributed:
        # single broadcast of every index array from rank 0 over NCCL / NVLink
        from bowtie2_b200.dist import broadcast_index
        torch.cuda.synchronize(); dist.barrier()
        tb = time.time()
        desc, tensors = broadcast_index(built, 0, dev)
        torch.cuda.synchronize(); dist.barrier()
        bcast_s = time.time() - tb
        gpu.load_index_device(desc, keep=tensors)
    else:
        gpu.load_index_device(built.device_desc(dev), keep=built)
    info = gpu.info()
    ktab_s = 0.0
    if args.seed_table > info["ftab_chars"]:
        torch.cuda.synchronize(); tk = time.time()
        gpu.build_seed_table(args.seed_table)
        ktab_s = time.time() - tk
        log(f"rank {rank}: {args.seed_table}-mer seed table built in {ktab_s:.2f}s")
    sa_s = 0.0
    if 0 <= args.dense_sa < info["off_rate"]:
        torch.cuda.synchronize(); tk = time.time()
        gpu.build_dense_sa(args.dense_sa)
        sa_s = time.time() - tk
        log(f"rank {rank}: SA sample of rate {args.dense_sa} built in {sa_s:.2f}s")
    if paired:
        reads, quals = make_pairs_gpu(torch, dev, contigs, args.reads, READ_LEN, seed=1 + rank)
    else:
        reads, quals = make_reads_gpu(torch, dev, contigs, args.reads, READ_LEN, seed=1 + rank)
    index_base = os.path.join(WORKDIR, "idx")
    if need_files:
        os.makedirs(WORKDIR, exist_ok=True)
        built.write_files(index_base)
        log(f"index files written to {index_base}.*.bt2")
    del contigs
    torch.cuda.synchronize()
    log(f"rank {rank}: setup {time.time() - t0:.1f}s, index {info['device_bytes'] / 1e9:.2f} GB in HBM")

    B = min(args.batch, args.reads)          # units (pairs / reads) per step
    nb = args.reads // B
    BR = B * mates                            # reads per step
    offs = torch.arange(0, (BR + 1) * READ_LEN, READ_LEN, dtype=torch.int64, device=dev)

    # ---- CPU baseline / reference arm -----------------------------------------------------------
    cpu_baseline = None
    if need_files:
        n_big = args.cpu_sample or int(min(args.reads, 400_000 if paired else 1_000_000))
        n_small = max(n_big // 10, 1000)
        r_np = reads[:n_big * mates].cpu().numpy(); q_np = quals[:n_big * mates].cpu().numpy()
        if paired:
            fq_big = (os.path.join(WORKDIR, "big_1.fq"), os.path.join(WORKDIR, "big_2.fq"))
            fq_small = (os.path.join(WORKDIR, "small_1.fq"), os.path.join(WORKDIR, "small_2.fq"))
            for m in range(2):
                write_fastq(fq_big[m], r_np[m::2], q_np[m::2])
                write_fastq(fq_small[m], r_np[m:2 * n_small:2], q_np[m:2 * n_small:2])
        else:
            fq_big, fq_small = os.path.join(WORKDIR, "big.fq"), os.path.join(WORKDIR, "small.fq")
            write_fastq(fq_big, r_np, q_np)
            write_fastq(fq_small, r_np[:n_small], q_np[:n_small])
        del r_np, q_np
        threads = best_thread_count(index_base, fq_small, n_small, cores, ref_preset) if os.path.exists(ref_binary()[0]) else cores
        sample = (f"{{}} {unit}: difference of a {n_big}- and a {n_small}-{unit[:-1]} run of {{}} {' '.join(ref_preset)} -p {threads} "
                  f"(fastest thread count of those tried on {cores} hardware threads; index load cancels{{}})")
        if args.impl == "reference":
            per = []
            for s in range(args.warmup + args.steps):
                r = time_reference(index_base, fq_small, fq_big, n_small, n_big, threads, ref_preset)
                if r is None:
                    print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bowtie2-align-s not built"}))
                    return 0
                if s >= args.warmup:
                    per.append(r)
                if s == 0 and r["t_big"] > 60:      # keep the whole run within a few minutes
                    per = per or [r]
                    break
            rps = float(np.median([p["reads_per_s"] for p in per]))
            val = rps / 1e6
            line = {"metric": "Mreads/s", "value": val, "unit": "Mreads/s", "n_gpus": 0, "steps": len(per), "warmup": args.warmup,
                    "ms_per_step": 1e3 * (n_big - n_small) / rps, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "u8/i16 (SSE/AVX2 striped DP), u64 popcount FM", "data": "synthetic",
                    "impl": "reference",
                    "config": {"workload": workload, "full_size": full, "read_unit": unit[:-1], "preset": " ".join(ref_preset),
                               "host_threads": cores, "cgroup_cpu_quota": cpu_quota},
                    "cpu_baseline": {"value": val, "unit": "Mreads/s", "cores": threads, "kind": "reference",
                                     "sample": sample.format(n_big - n_small, per[0]["binary"], "")},
                    "e2e": {"value": val, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        if rank == 0 and not args.no_cpu_baseline:
            r = time_reference(index_base, fq_small, fq_big, n_small, n_big, threads, ref_preset)
            if r is not None:
                cpu_baseline = {"value": r["reads_per_s"] / 1e6, "unit": "Mreads/s", "cores": threads, "kind": "reference",
                                "sample": sample.format(n_big - n_small, r["binary"], f"; {r['t_big']:.1f}s and {r['t_small']:.1f}s wall")}
            log("cpu baseline:", cpu_baseline)
        shutil.rmtree(WORKDIR, ignore_errors=True)

    # ---- our arm -----------------------------------------------------------------------------------
    pipe = Pipeline(gpu, wl["preset"], max_len=READ_LEN, max_reads=BR, row_cap=16, range_max=8, max_cands=48, max_alns=2,
                    max_probs=4 * B, both_mates=paired)
    if paired:
        pipe.enable_pairs()
    stream = torch.cuda.Stream(device=dev)       # explicit non-default stream: kernels and timing events share it
    torch.cuda.set_stream(stream)

    def step_dev(i, count=False):
        k = i % nb
        r, q = reads[k * BR:(k + 1) * BR], quals[k * BR:(k + 1) * BR]
        if paired:
            pipe.run_paired_dev(r.data_ptr(), q.data_ptr(), offs.data_ptr(), B, stream=stream.cuda_stream, count=count)
        else:
            pipe.run_dev(r.data_ptr(), q.data_ptr(), offs.data_ptr(), B, stream=stream.cuda_stream, count=count)

    def stages():
        s = pipe.stage_ms()
        if paired:
            s.update(pipe.pair_stage_ms())
        return s

    # counters (algorithmic work) from one untimed counting pass
    step_dev(0, count=True)
    torch.cuda.synchronize()
    cnt = pipe.counters()
    if paired:
        cnt.update(pipe.pair_counters())
    for i in range(args.warmup):
        step_dev(i)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for i in range(args.steps):
        step_dev(args.warmup + i)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    # per-stage times: re-run the steps once more, reading the stage events after each (outside the headline timing)
    names = list(stages().keys())
    stage_acc = np.zeros(len(names))
    for i in range(args.steps):
        step_dev(args.warmup + i)
        s = stages()
        stage_acc += np.array([s[k] for k in names])
    stage_ms = dict(zip(names, (stage_acc / args.steps).tolist()))
    # ---- e2e: host buffers through the C ABI (H2D of the batch + D2H of results inside the timed region)
    import ctypes as C
    nbuf = min(2, nb)
    hseq = [torch.empty(BR * READ_LEN, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    hqual = [torch.empty(BR * READ_LEN, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    for k in range(nbuf):
        hseq[k].copy_(reads[k * BR:(k + 1) * BR].reshape(-1)); hqual[k].copy_(quals[k * BR:(k + 1) * BR].reshape(-1))
    hoff = np.arange(0, (BR + 1) * READ_LEN, READ_LEN, dtype=np.uint64)
    hres = torch.empty(BR * READ_RESULT.itemsize, dtype=torch.uint8).pin_memory()
    hops = torch.empty(BR * pipe.max_ops, dtype=torch.uint8).pin_memory()
    hpairs = torch.empty(max(B, 1) * PAIR_RESULT.itemsize, dtype=torch.uint8).pin_memory()

    def step_host(i):
        k = i % nbuf
        st_ = _Reads(BR, hseq[k].data_ptr(), hqual[k].data_ptr(), hoff.ctypes.data)
        if paired:
            gpu._check(gpu._lib.bt2g_pipeline_run_paired_host(pipe._h, C.byref(st_), hres.data_ptr(), hops.data_ptr(), hpairs.data_ptr()),
                       "bt2g_pipeline_run_paired_host")
        else:
            gpu._check(gpu._lib.bt2g_pipeline_run_host(pipe._h, C.byref(st_), hres.data_ptr(), hops.data_ptr()), "bt2g_pipeline_run_host")

    for i in range(min(args.warmup, 3)):
        step_host(i)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    t_e2e0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t_e2e0
    res_np = np.frombuffer(hres.numpy().tobytes(), dtype=READ_RESULT)
    found = float((res_np["found"] & 0xff != 0).mean())
    overflow = int((res_np["found"] & 0x100 != 0).sum())
    conc = None
    if paired:
        pr = np.frombuffer(hpairs.numpy().tobytes(), dtype=PAIR_RESULT)
        conc = {"concordant_frac": float((pr["pair_type"] == 1).mean()), "by_mate_dp_frac": float((pr["source"] != 0).mean())}
    clk = clocks.stop() if rank == 0 else None
    # informational: host-side SAM formatting rate of the last batch's results (bt2g_sam_format on the host threads the
    # container may use); not part of `value` or `e2e` -- the records/s it sustains is the next bottleneck
    sam_info = None
    if rank == 0:
        try:
            from bowtie2_b200.lib import ReadBatch, sam_format
            nfmt = min(BR, 200_000) // mates * mates
            kl = (args.steps - 1) % nbuf                    # the host buffers of the last timed step
            rb = ReadBatch(hseq[kl].numpy()[:nfmt * READ_LEN], hoff[:nfmt + 1], hqual[kl].numpy()[:nfmt * READ_LEN])
            res_f = np.frombuffer(hres.numpy().tobytes(), dtype=READ_RESULT)[:nfmt]
            ops_f = hops.numpy()[:nfmt * pipe.max_ops].reshape(nfmt, pipe.max_ops)
            prs_f = np.frombuffer(hpairs.numpy().tobytes(), dtype=PAIR_RESULT)[:nfmt // 2] if paired else None
            thr = int(cpu_quota) if cpu_quota else min(cores, 16)
            t_f = time.perf_counter()
            txt = sam_format(gpu._lib, rb, res_f, ops_f, [f"chr{k + 1}" for k in range(GENOME_CONTIGS)], pairs=prs_f, threads=max(thr, 1),
                             as_bytes=True)
            dt_f = time.perf_counter() - t_f
            sam_info = {"records": nfmt, "threads": max(thr, 1), "Mrecords_per_s": nfmt / dt_f / 1e6, "bytes_per_record": len(txt) / max(nfmt, 1),
                        "note": "host formatter (one pass, multi-threaded); device-side formatting is next"}
        except Exception as e:                      # never let the informational extra break the bench line
            sam_info = {"error": repr(e)[:200]}

    # max over ranks
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    total_units = args.steps * B * world
    value = total_units / (ms_max / 1e3) / 1e6
    e2e_val = total_units / (e2e_ms_max / 1e3) / 1e6

    if rank == 0:
        side = info["side_sz"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the stage with the largest device time
        alg_bytes = {"exact_sweep": cnt["sweep_sides"] * side, "seed_search": cnt["seed_sides"] * side,
                     "resolve": cnt["resolve_sides"] * side, "dp": cnt["dp_cells"] * 1}
        if paired:
            alg_bytes["mate_dp"] = cnt["mate_cells"] * 1
        dom = max(alg_bytes, key=lambda k: stage_ms[k])
        fm_bytes = (cnt["sweep_sides"] + cnt["seed_sides"] + cnt["resolve_sides"]) * side
        fm_ms = stage_ms["exact_sweep"] + stage_ms["seed_search"] + stage_ms["resolve"]
        dp_ms = stage_ms["dp"] + (stage_ms["mate_dp"] if paired else 0.0)
        dp_cells = cnt["dp_cells"] + (cnt["mate_cells"] if paired else 0)
        roof = {"bound": "hbm", "kernel": dom, "achieved": alg_bytes[dom] / (stage_ms[dom] / 1e3) / 1e9, "peak": peak,
                "unit": "GB/s", "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": alg_bytes[dom], "kernel_ms": stage_ms[dom],
                "fm_stages": {"achieved": fm_bytes / (fm_ms / 1e3) / 1e9, "frac": fm_bytes / (fm_ms / 1e3) / 1e9 / peak,
                              "bytes_per_read": fm_bytes / cnt["reads"]},
                "dp_gcups": dp_cells / (dp_ms / 1e3) / 1e9}
        roof["frac"] = roof["achieved"] / peak
        pipeline_desc = ("exactSweep + multiseed round 0 + resolve(all rows of ranges<=8, cap 16) + DP/backtrace per distinct diagonal"
                         + ("; then mate framing + mate-finding DP for anchors without a concordant independent mate + pair pick" if paired else ""))
        line = {"metric": "Mreads/s", "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 popcount (FM rank) + s16x2 DPX (DP)", "data": "synthetic",
                "config": {"workload": workload, "full_size": full, "read_unit": unit[:-1], "mates_per_s_M": value * mates,
                           "batch": B, "preset": " ".join(ref_preset),
                           "l2": "inputs larger than L2 (random access over a %.1f GB index; a different batch each step)" % (info["device_bytes"] / 1e9),
                           "pipeline": pipeline_desc, "seed_table_k": args.seed_table, "seed_table_build_s": ktab_s, "dense_sa_rate": args.dense_sa, "dense_sa_build_s": sa_s, "index_bcast_s": bcast_s, "aligned_frac": found, "pairs": conc,
                           "host_threads": cores, "cgroup_cpu_quota": cpu_quota, "dp_workspace_overflows": overflow},
                "clocks": clk, "gpu_launches": pipe.kernel_launches() * args.steps,
                "e2e": {"value": e2e_val, "unit": "Mreads/s", "h2d_bytes_per_step": 2 * BR * READ_LEN + (BR + 1) * 8,
                        "d2h_bytes_per_step": BR * READ_RESULT.itemsize + BR * pipe.max_ops + (B * PAIR_RESULT.itemsize if paired else 0)},
                "roofline": roof, "stage_ms": stage_ms, "work_per_step": cnt, "sam_format_host": sam_info, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
