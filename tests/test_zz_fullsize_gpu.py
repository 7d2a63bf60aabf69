"""GPU, BASELINE.json full size (3 Gbp .bt2 index built on the device): properties that need no oracle.

The oracle-based parity tests run on small indexes (the reference and its restatement finish in seconds there); at
the size the headline metric is quoted on, correctness is checked through what the domain guarantees:
  * round trip: a read copied from the genome aligns back, exactly, to a locus whose reference text equals it
    (the sampled locus itself unless it lies in a repeat family);
  * strand symmetry: the reverse complement of a read lands on the same locus with the opposite strand;
  * idempotence: the same batch twice gives identical result arrays;
  * score checksum: a read with k substitutions at quality 40 reports score -6 k and exactly k mismatch ops,
    no gaps, at its origin;
  * invariance: the extended seed table and the dense SA sample change no result byte.
This file sorts last so that `pytest -x` has run every other test before it."""
import os
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500)]

N_CONTIGS, CONTIG_LEN, RDLEN, NREADS = 24, 125_000_000, 150, 120_000


def _comp_rev(torch, x):
    comp = torch.tensor([3, 2, 1, 0, 4], dtype=torch.uint8, device=x.device)
    return comp[x.flip(1).long()]


def test_fullsize_properties():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    if torch.cuda.get_device_properties(0).total_memory < 150 * (1 << 30):
        pytest.skip("needs a 180 GB class device for the 3 Gbp index build")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from bowtie2_b200 import Bt2Gpu
    from bowtie2_b200.index_build import build_index
    from bowtie2_b200.lib import OP_MATCH, OP_MM, Pipeline, ReadBatch

    contigs = bench.make_genome_gpu(torch, dev, N_CONTIGS, CONTIG_LEN)
    built = build_index(contigs)
    torch.cuda.synchronize()
    gpu = Bt2Gpu(0)
    gpu.load_index_device(built.device_desc(dev), keep=built)
    assert gpu.info()["len"] > 2_900_000_000

    # reads = windows of the genome, away from the N gap in the middle of each contig; half of them reverse-complemented
    g = torch.Generator(device=dev)
    g.manual_seed(4711)
    ci = torch.randint(0, N_CONTIGS, (NREADS,), device=dev, generator=g)
    half = CONTIG_LEN // 2
    pos = torch.randint(1000, half - 21_000, (NREADS,), device=dev, generator=g)      # read ends stay inside the contig
    pos = torch.where(torch.rand(NREADS, device=dev, generator=g) < 0.5, pos, pos + half + 20_000)
    genome = torch.cat(contigs)
    ar = torch.arange(RDLEN, device=dev)
    fwd = genome[(ci * CONTIG_LEN + pos)[:, None] + ar[None, :]]
    assert int((fwd > 3).sum()) == 0
    rc = torch.rand(NREADS, device=dev, generator=g) < 0.5
    reads = torch.where(rc[:, None], _comp_rev(torch, fwd), fwd)
    off = np.arange(0, (NREADS + 1) * RDLEN, RDLEN, dtype=np.uint64)
    quals = np.full(NREADS * RDLEN, ord("I"), dtype=np.uint8)

    def run(r):
        batch = ReadBatch(r.reshape(-1).cpu().numpy(), off, quals)
        return pipe.run_host(batch)

    def text_at(res):
        """reference text under every reported alignment start, read-length long, in read orientation"""
        t = torch.from_numpy(res["tidx"].astype(np.int64)).to(dev)
        o = torch.from_numpy(res["refoff"].astype(np.int64)).to(dev)
        w = genome[(t * CONTIG_LEN + o)[:, None] + ar[None, :]]
        f = torch.from_numpy(res["fw"].astype(np.int64)).to(dev) != 0
        return torch.where(f[:, None], w, _comp_rev(torch, w))

    pipe = Pipeline(gpu, "sensitive", max_len=RDLEN, max_reads=NREADS, row_cap=16, range_max=8)
    try:
        # ---- round trip on exact copies
        res, ops = run(reads)
        assert ((res["found"] & 0xff) == 2).all()                 # every copy is an exact end-to-end hit
        assert bool((text_at(res) == reads).all())                # ... at a locus that spells the read
        truth_t, truth_o, truth_fw = ci.cpu().numpy(), pos.cpu().numpy(), (~rc).cpu().numpy()
        at_origin = (res["tidx"] == truth_t) & (res["refoff"] == truth_o) & ((res["fw"] != 0) == truth_fw)
        assert at_origin.mean() > 0.97                             # the rest sit in repeat families (another exact copy)
        # ---- idempotence
        res2, ops2 = run(reads)
        assert np.array_equal(res, res2) and np.array_equal(ops, ops2)
        # ---- strand symmetry
        res3, _ = run(_comp_rev(torch, reads))
        same_locus = (res3["tidx"] == res["tidx"]) & (res3["refoff"] == res["refoff"]) & ((res3["fw"] != 0) != (res["fw"] != 0))
        assert same_locus[at_origin].mean() > 0.999
        # ---- k substitutions: score checksum through seed search + DP
        k = 2
        mut = reads.clone()
        cols = torch.stack([torch.randint(10, 70, (NREADS,), device=dev, generator=g),
                            torch.randint(80, 140, (NREADS,), device=dev, generator=g)], dim=1)
        rows = torch.arange(NREADS, device=dev)
        for j in range(k):
            c = cols[:, j]
            mut[rows, c] = (mut[rows, c] + 1 + torch.randint(0, 3, (NREADS,), device=dev, generator=g).to(torch.uint8)) % 4
        resm, opsm = run(mut)
        aligned = (resm["found"] & 0xff) == 1
        # reads that lie wholly inside a 120-copy repeat family (1 % of the genome) have no seed range under range_max
        assert aligned.mean() > 0.97
        origin_m = aligned & (resm["tidx"] == truth_t) & (resm["refoff"] == truth_o) & ((resm["fw"] != 0) == truth_fw)
        assert origin_m.mean() > 0.95
        sel = np.nonzero(origin_m)[0]
        assert (resm["score"][sel] == -6 * k).all() and (resm["nops"][sel] == RDLEN).all()
        typ = opsm[sel][:, :RDLEN] & 3
        assert ((typ == OP_MATCH) | (typ == OP_MM)).all() and ((typ == OP_MM).sum(axis=1) == k).all()
        # ---- invariance under the load-time acceleration structures
        gpu.build_seed_table(14)
        gpu.build_dense_sa(2)
        try:
            resa, opsa = run(mut)
            rese, opse = run(reads)
        finally:
            gpu.build_seed_table(0)
            gpu.build_dense_sa(-1)
        def same_ops(a, na, b):                            # op bytes are defined up to nops (exact end-to-end hits have none)
            cols = np.arange(a.shape[1])[None, :]
            live = cols < na[:, None]
            return np.array_equal(np.where(live, a, 0), np.where(live, b, 0))
        assert np.array_equal(resa, resm) and same_ops(opsa, resa["nops"], opsm)
        assert np.array_equal(rese, res) and same_ops(opse, rese["nops"], ops)
    finally:
        pipe.close()
        del genome, contigs, built
        torch.cuda.empty_cache()
