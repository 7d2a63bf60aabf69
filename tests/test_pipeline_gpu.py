"""GPU: the batched hot path (bt2g_pipeline) end to end.

(1) stage consistency: the pipeline's per-read result equals what the stand-alone, individually
    parity-tested entry points give for the same read;
(2) against the unmodified reference run as a program (oracle/_ref/bowtie2-align-s): for reads
    the reference aligns, the pipeline finds the same locus / strand / score / CIGAR.  The
    pipeline resolves candidates speculatively instead of replaying the reference's sequential,
    RNG-driven policy, so this is a high-concordance property, not bit-identity."""
import os
import subprocess

import numpy as np
import pytest

from bowtie2_b200 import policy, synth
from bowtie2_b200.lib import Pipeline, ReadBatch, ops_to_cigar
from oracle_lib import have_reference, ref_bin

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _norm_sam(line):
    """a SAM record without the two fields that depend on the sequential search policy (MAPQ, XS:i)"""
    f = line.split("\t")
    f[4] = "."
    return "\t".join(x for x in f if not x.startswith("XS:i:"))


def _run_reference(index, fq, preset="--sensitive", mode="--end-to-end"):
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), preset, mode, "--seed", "0", "-p", "4", "--reorder",
                                   "-x", index, "-U", fq], stderr=subprocess.DEVNULL).decode()
    recs = []
    for line in out.splitlines():
        if line.startswith("@"):
            continue
        f = line.split("\t")
        tags = {t[:2]: t[5:] for t in f[11:]}
        recs.append(dict(flag=int(f[1]), rname=f[2], pos=int(f[3]) - 1, mapq=int(f[4]), cigar=f[5],
                         AS=int(tags["AS"]) if "AS" in tags else None, XS=int(tags["XS"]) if "XS" in tags else None, line=line))
    return recs


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("rdlen", [100, 150])
def test_pipeline_vs_reference_program(gpu, synth_index, synth_genome, tmp_path, rdlen):
    gpu.load_index_files(synth_index)
    reads, quals, truth = synth.make_reads(synth_genome, 3000, rdlen, seed=77 + rdlen, sub_rate=0.01, indel_rate=0.001)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, reads, quals)
    want = _run_reference(synth_index, fq)
    pipe = Pipeline(gpu, "sensitive", max_len=rdlen, max_reads=4096, row_cap=16, range_max=16)
    res, ops = pipe.run_host(ReadBatch.from_list(reads, quals))
    n_ref_aln = n_same = n_same_cigar = n_unique = n_unique_same = n_gpu_only = n_same_mapq = 0
    for i, w in enumerate(want):
        r = res[i]
        if w["flag"] & 4:
            n_gpu_only += r["found"] != 0
            continue
        n_ref_aln += 1
        same = (r["found"] != 0 and int(r["tidx"]) == int(w["rname"][3:]) - 1 and int(r["refoff"]) == w["pos"]
                and bool(r["fw"]) == (not (w["flag"] & 16)) and int(r["score"]) == w["AS"])
        n_same += same
        if w["mapq"] >= 30:
            n_unique += 1
            n_unique_same += same
        if same:
            cig = f"{rdlen}M" if r["found"] == 2 else ops_to_cigar(ops[i], int(r["nops"]))
            n_same_cigar += cig == w["cigar"]
            n_same_mapq += int(r["mapq"]) == w["mapq"]
    # reads in -> SAM out: records formatted from the pipeline's results equal the reference program's lines
    from bowtie2_b200.lib import load_library, sam_format
    names = [f"r{i}" for i in range(len(reads))]
    text = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, [f"chr{k + 1}" for k in range(len(synth_genome))],
                      read_names=names, threads=2).rstrip("\n").split("\n")
    n_line = n_line_same = 0
    for i, w in enumerate(want):
        r = res[i]
        if w["flag"] & 4 or not (r["found"] != 0 and int(r["refoff"]) == w["pos"] and int(r["score"]) == w["AS"]
                                 and bool(r["fw"]) == (not (w["flag"] & 16)) and int(r["tidx"]) == int(w["rname"][3:]) - 1):
            continue
        n_line += 1
        n_line_same += _norm_sam(text[i]) == _norm_sam(w["line"])
    assert n_line_same >= 0.99 * n_line, (n_line_same, n_line)
    assert n_ref_aln > 2500
    # the MAPQ formula is the reference's; the runner-up score feeding it comes from the speculative pipeline
    assert n_same_mapq >= 0.9 * n_same, (n_same_mapq, n_same)
    # reads the reference places confidently must agree; repeats may legitimately differ
    assert n_unique_same >= 0.995 * n_unique, (n_unique_same, n_unique)
    assert n_same >= 0.97 * n_ref_aln, (n_same, n_ref_aln)
    assert n_same_cigar >= 0.995 * n_same, (n_same_cigar, n_same)
    pipe.close()


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_pipeline_local_vs_reference_program(gpu, synth_index, synth_genome, tmp_path):
    """--local --sensitive-local: soft-clipped alignments (reads carry 12 random bases at one end)."""
    gpu.load_index_files(synth_index)
    rdlen = 120
    reads, quals, truth = synth.make_reads(synth_genome, 2000, rdlen, seed=311, sub_rate=0.01, indel_rate=0.001)
    rng = np.random.default_rng(9)
    for k, r in enumerate(reads):
        if k % 2 == 0:
            junk = rng.integers(0, 4, size=12)
            if k % 4 == 0:
                r[:12] = junk
            else:
                r[-12:] = junk
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, reads, quals)
    want = _run_reference(synth_index, fq, preset="--sensitive-local", mode="--local")
    pipe = Pipeline(gpu, "sensitive", max_len=rdlen, max_reads=2048, row_cap=16, range_max=16, local=True, max_cands=256)
    res, ops = pipe.run_host(ReadBatch.from_list(reads, quals))
    n_ref_aln = n_same = n_same_cigar = n_unique = n_unique_same = n_clipped = 0
    for i, w in enumerate(want):
        r = res[i]
        if w["flag"] & 4:
            continue
        n_ref_aln += 1
        same = (r["found"] != 0 and int(r["tidx"]) == int(w["rname"][3:]) - 1 and int(r["refoff"]) == w["pos"]
                and bool(r["fw"]) == (not (w["flag"] & 16)) and int(r["score"]) == w["AS"])
        n_same += same
        if w["mapq"] >= 30:
            n_unique += 1
            n_unique_same += same
        if same:
            cig = f"{rdlen}M" if r["found"] == 2 else ops_to_cigar(ops[i], int(r["nops"]), int(r["trim_left"]), int(r["trim_right"]))
            n_same_cigar += cig == w["cigar"]
            n_clipped += "S" in cig
    assert n_ref_aln > 1800
    assert n_clipped > 500
    assert n_unique_same >= 0.99 * n_unique, (n_unique_same, n_unique)
    assert n_same >= 0.97 * n_ref_aln, (n_same, n_ref_aln)
    assert n_same_cigar >= 0.99 * n_same, (n_same_cigar, n_same)
    pipe.close()
    gpu.set_scoring(local=False)


def test_pipeline_stage_consistency(gpu, synth_index, synth_genome):
    gpu.load_index_files(synth_index)
    rdlen = 100
    reads, quals, truth = synth.make_reads(synth_genome, 400, rdlen, seed=5, sub_rate=0.02, indel_rate=0.003)
    batch = ReadBatch.from_list(reads, quals)
    pipe = Pipeline(gpu, "sensitive", max_len=rdlen, max_reads=1024, row_cap=16, range_max=16)
    res, ops = pipe.run_host(batch)
    # same answer through the device-pointer entry point, and counters are populated
    import torch
    dseq = torch.from_numpy(batch.seq).cuda(); dq = torch.from_numpy(batch.qual).cuda()
    doff = torch.from_numpy(batch.off.astype(np.int64)).cuda()
    pipe.run_dev(dseq.data_ptr(), dq.data_ptr(), doff.data_ptr(), batch.n, count=True)
    torch.cuda.synchronize()
    c = pipe.counters()
    assert c["reads"] == batch.n and c["sweep_sides"] > 0 and c["seed_sides"] > 0 and c["dp_cells"] > 0
    res2, _ = pipe.run_host(batch)
    assert np.array_equal(res, res2)
    # the chunked host path (copies overlapped with compute) gives the same answers
    os.environ["BT2G_HOST_CHUNK_MIN"] = "64"
    try:
        res3, ops3 = pipe.run_host(batch)
    finally:
        del os.environ["BT2G_HOST_CHUNK_MIN"]
    assert np.array_equal(res, res3)
    for i in range(batch.n):
        assert np.array_equal(ops[i, :res[i]["nops"]], ops3[i, :res3[i]["nops"]])
    # exact end-to-end hits are exactly the reads whose exact sweep reports a range
    mine, ee = gpu.exact_sweep(batch)
    has_ee = (ee[:, 1] > ee[:, 0]) | (ee[:, 3] > ee[:, 2])
    assert np.array_equal(res["found"] == 2, has_ee)
    assert (res["found"] != 0).sum() > 350
    # MAPQ = BowtieMapq2 on (best, runner-up) with the read's minimum and perfect scores
    sc = policy.Scoring.default(False)
    INT_MIN = -(1 << 31)
    for i in range(batch.n):
        r = res[i]
        if r["found"] == 0:
            continue
        sec = None if int(r["score2"]) == INT_MIN else int(r["score2"])
        assert int(r["mapq"]) == policy.mapq_v2(int(r["score"]), sec, sc.min_score(rdlen), sc.perfect_score(rdlen), True), i
    pipe.close()


def _write_pair_fastq(tmp_path, reads, quals):
    f1, f2 = str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq")
    synth.write_fastq(f1, reads[0::2], quals[0::2])
    synth.write_fastq(f2, reads[1::2], quals[1::2])
    return f1, f2


def _run_reference_paired(index, f1, f2, preset="--sensitive"):
    out = subprocess.check_output([ref_bin("bowtie2-align-s"), preset, "--end-to-end", "--seed", "0", "-p", "4", "--reorder",
                                   "-x", index, "-1", f1, "-2", f2], stderr=subprocess.DEVNULL).decode()
    recs = []
    for line in out.splitlines():
        if line.startswith("@"):
            continue
        f = line.split("\t")
        tags = {t[:2]: t[5:] for t in f[11:]}
        recs.append(dict(flag=int(f[1]), rname=f[2], pos=int(f[3]) - 1, mapq=int(f[4]), cigar=f[5], tlen=int(f[8]),
                         AS=int(tags["AS"]) if "AS" in tags else None, YT=tags.get("YT")))
    return recs


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("rdlen,preset,tables", [(100, "sensitive", False), (150, "very-sensitive", True)])
def test_paired_pipeline_vs_reference_program(gpu, synth_index, synth_genome, tmp_path, rdlen, preset, tables):
    """FR pairs, -I 0 -X 500: concordant pairs of the reference are found with the same placement,
    including pairs where one mate has no exact seed and only the mate-finding DP can place it.
    The second case is the headline configuration (2x150, --very-sensitive) with the seed table and the dense SA on."""
    gpu.load_index_files(synth_index)
    if tables:
        gpu.build_seed_table(12)
        gpu.build_dense_sa(0)
    npairs = 1500
    reads, quals, truth = synth.make_pairs(synth_genome, npairs, rdlen, seed=4242 + rdlen, sub_rate=0.01, indel_rate=0.001,
                                           hard_frac=0.25, hard_period=12 if rdlen == 100 else 14)
    f1, f2 = _write_pair_fastq(tmp_path, reads, quals)
    want = _run_reference_paired(synth_index, f1, f2, preset="--" + preset)
    assert len(want) == 2 * npairs
    pipe = Pipeline(gpu, preset, max_len=rdlen, max_reads=4096, row_cap=16, range_max=16, both_mates=True)
    pipe.enable_pairs()
    res, ops, pairs = pipe.run_paired_host(ReadBatch.from_list(reads, quals))
    n_cp = n_cp_same = n_rescued = n_rescued_same = n_cigar = 0
    for i in range(npairs):
        w1, w2 = want[2 * i], want[2 * i + 1]
        if w1["YT"] != "CP" or w2["YT"] != "CP":
            continue
        n_cp += 1
        ok = pairs[i]["pair_type"] == 1
        for k, w in ((2 * i, w1), (2 * i + 1, w2)):
            r = res[k]
            ok = ok and (r["found"] & 0xff) != 0 and int(r["tidx"]) == int(w["rname"][3:]) - 1 and int(r["refoff"]) == w["pos"] \
                and bool(r["fw"]) == (not (w["flag"] & 16)) and int(r["score"]) == w["AS"]
        n_cp_same += ok
        if ok:
            assert int(pairs[i]["fraglen"]) == abs(w1["tlen"])
            for k, w in ((2 * i, w1), (2 * i + 1, w2)):
                cig = f"{rdlen}M" if (res[k]["found"] & 0xff) == 2 else ops_to_cigar(ops[k], int(res[k]["nops"]))
                n_cigar += cig == w["cigar"]
        if pairs[i]["source"] != 0:
            n_rescued += 1
            n_rescued_same += ok
    assert n_cp > 0.9 * npairs
    assert n_cp_same >= 0.97 * n_cp, (n_cp_same, n_cp)
    assert n_cigar >= 0.99 * 2 * n_cp_same, (n_cigar, n_cp_same)
    assert n_rescued > 100, n_rescued                      # the mate DP really ran and decided pairs
    assert n_rescued_same >= 0.9 * n_rescued, (n_rescued_same, n_rescued)
    # chunked host path (copies overlapped with compute): identical answers
    os.environ["BT2G_HOST_CHUNK_MIN"] = "64"
    try:
        res3, ops3, pairs3 = pipe.run_paired_host(ReadBatch.from_list(reads, quals))
    finally:
        del os.environ["BT2G_HOST_CHUNK_MIN"]
    assert np.array_equal(res, res3) and np.array_equal(pairs, pairs3)
    pipe.close()
    if tables:
        gpu.build_seed_table(0)
        gpu.build_dense_sa(-1)
