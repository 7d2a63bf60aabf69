"""GPU: the exact search policy over the library's own entry points (byte-identical SAM to the reference program), the DP
kernels' per-candidate fates against the oracle's attempt log, whole files through align_files, and the --offrate override.
All through the C ABI (ctypes)."""
import os
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("which", ["synth_index", "synth_index_large"])
def test_offrate_override_changes_no_offset(which, request):
    """--offrate larger than the index's own (bt2_io.cpp:217-230): the sparser SA sample lengthens the walk of
    getOffset and changes no resolved offset; checked against the oracle's walk on the full sample."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bowtie2_b200 import Bt2Gpu
    from oracle_lib import Oracle
    base = request.getfixturevalue(which)
    O = Oracle(base)
    sc = O.scalars()
    rng = np.random.default_rng(26)
    rows = np.concatenate([rng.integers(0, sc["bwt_len"], 3000), [sc["z_off"], 0, sc["bwt_len"] - 1]]).astype(np.uint64)
    hitlen = rng.integers(1, 60, len(rows)).astype(np.uint32)
    want = O.get_offset(rows)
    for extra in (0, 1, 3):
        g = Bt2Gpu(0)
        g.load_index_files(base, offrate=sc["off_rate"] + extra)
        assert g.info()["off_rate"] == sc["off_rate"] + extra
        assert np.array_equal(g.resolve(rows, hitlen, False)[0], want)
        g.close()


def _records(path):
    return [l.rstrip("\n") for l in open(path) if not l.startswith("@")]


def _norm(line):
    f = line.split("\t")
    f[4] = "."
    return "\t".join(x for x in f if not x.startswith("XS:i:"))


@pytest.mark.parametrize("paired", [False, True])
def test_files_in_sam_out(paired, lambda_index, tmp_path):
    """bowtie2_b200.align.align_files on the golden lambda reads: header identical to the reference program's, records in
    input order, and the records of reads placed at the reference's locus identical apart from MAPQ / XS:i."""
    import io
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bowtie2_b200.align import align_files
    from conftest import GOLDEN
    golden_path = os.path.join(GOLDEN, "lambda_P_sensitive.sam" if paired else "lambda_U_sensitive.sam")
    out = str(tmp_path / "out.sam")
    r1, r2 = os.path.join(GOLDEN, "lambda_reads_1.fq"), os.path.join(GOLDEN, "lambda_reads_2.fq")
    summ = io.StringIO()
    if paired:
        # the golden holds the first 200 pairs
        for src, dst in ((r1, "a.fq"), (r2, "b.fq")):
            with open(src) as f, open(tmp_path / dst, "w") as g:
                g.writelines(f.readlines()[:800])
        counts = align_files(lambda_index, out, str(tmp_path / "a.fq"), str(tmp_path / "b.fq"), batch_reads=128, threads=2, summary=summ)
    else:
        counts = align_files(lambda_index, out, r1, batch_reads=700, threads=2, summary=summ)
    want_hdr = [l for l in open(golden_path) if l.startswith("@")]
    got_hdr = [l for l in open(out) if l.startswith("@")]
    assert got_hdr == want_hdr
    want, got = _records(golden_path), _records(out)
    assert len(got) == len(want)
    assert [l.split("\t")[0] for l in got] == [l.split("\t")[0] for l in want]
    n_al = n_locus = n_same = 0
    for g, w in zip(got, want):
        fw, fg = w.split("\t"), g.split("\t")
        if int(fw[1]) & 4:
            continue
        n_al += 1
        if fg[1:4] == fw[1:4] and fg[5] == fw[5]:
            n_locus += 1
            n_same += _norm(g) == _norm(w)
    # the example reads are noisy (Ns, long indels, 30-250 bp); the speculative pipeline places most of them where the
    # reference does, and for those the whole record must agree
    assert n_locus >= 0.8 * n_al, (n_locus, n_al)
    # (a paired record also carries its mate's placement and score)
    assert n_same >= (0.85 if paired else 0.97) * n_locus, (n_same, n_locus)
    text = summ.getvalue()
    assert text.startswith(f"{len(want) // (2 if paired else 1)} reads; of these:\n") and text.endswith("overall alignment rate\n")
    assert int(counts["nread"][0]) == len(want) // (2 if paired else 1)


@pytest.mark.parametrize("paired", [False, True])
def test_exact_policy_over_gpu_primitives(paired, lambda_index, monkeypatch):
    """policy_engine over the GPU entry points (policy_backend_gpu.GpuBackend): byte-identical golden SAM, i.e. the
    reference's sequential policy with every hot-path primitive computed on the device."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bowtie2_b200 import Bt2Gpu
    from bowtie2_b200.lib import PAIR_RESULT, READ_RESULT, ReadBatch, load_library, sam_format
    from bowtie2_b200.policy_backend_gpu import GpuBackend
    from bowtie2_b200.policy_engine import PairedPolicyEngine, PolicyEngine
    from conftest import GOLDEN, read_fastq_codes
    from test_policy_engine import _fill
    g = Bt2Gpu(0)
    g.load_index_files(lambda_index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "lambda_P_sensitive.sam" if paired else "lambda_U_sensitive.sam"))
              if not l.startswith("@")]
    n = 100 if paired else 300
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_1.fq"), n)
    ref = ["gi|9626243|ref|NC_001416.1|"]
    if not paired:
        eng = PolicyEngine(GpuBackend(g), "sensitive")
        res = np.zeros(n, dtype=READ_RESULT)
        res["score2"] = -(1 << 31)
        ops = np.zeros((n, max(len(r) for r in r1) + 64), dtype=np.uint8)
        for i in range(n):
            r = eng.align_read(r1[i], q1[i], n1[i])
            if r.aligned:
                _fill(res, ops, i, r, r1[i])
        lines = sam_format(load_library(), ReadBatch.from_list(r1, q1), res, ops, ref, read_names=n1).rstrip("\n").split("\n")
        assert lines == golden[:n]
        # and in waves (one batched entry-point call per primitive and wave), through the whole-file driver
        import tempfile
        from bowtie2_b200.align import align_files
        out = os.path.join(tempfile.mkdtemp(), "exact.sam")
        align_files(lambda_index, out, os.path.join(GOLDEN, "lambda_reads_1.fq"), exact=True, batch_reads=1024, summary=None, gpu=g)
        got = [l.rstrip("\n") for l in open(out) if not l.startswith("@")]
        assert got == golden
    else:
        n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, "lambda_reads_2.fq"), n)
        il = lambda a, b: [x for p in zip(a, b) for x in p]
        reads, quals, names = il(r1, r2), il(q1, q2), il(n1, n2)
        eng = PairedPolicyEngine(GpuBackend(g), "sensitive")
        res = np.zeros(2 * n, dtype=READ_RESULT)
        res["score2"] = -(1 << 31)
        ops = np.zeros((2 * n, max(len(r) for r in reads) + 64), dtype=np.uint8)
        pairs = np.zeros(n, dtype=PAIR_RESULT)
        for i in range(n):
            pr = eng.align_pair(r1[i], q1[i], n1[i], r2[i], q2[i], n2[i])
            pairs[i]["pair_type"] = pr.pair_type
            for k in range(2):
                if pr.mates[k].aligned:
                    _fill(res, ops, 2 * i + k, pr.mates[k], reads[2 * i + k])
        lines = sam_format(load_library(), ReadBatch.from_list(reads, quals), res, ops, ref, read_names=names, pairs=pairs).rstrip("\n").split("\n")
        assert lines == golden[:2 * n]
        # the compiled engine in waves over the library's own entry points
        from bowtie2_b200.lib import policy_align, policy_backend_gpu, policy_params
        batch = ReadBatch.from_list(reads, quals)
        res2, ops2, pairs2, stats = policy_align(load_library(), policy_backend_gpu(g), policy_params("sensitive", paired=True, host_threads=4),
                                                 batch, names)
        lines2 = sam_format(load_library(), batch, res2, ops2, ref, read_names=names, pairs=pairs2).rstrip("\n").split("\n")
        assert lines2 == golden[:2 * n]
    g.close()


@pytest.mark.parametrize("local,cap", [(False, "1"), (True, "1"), (False, None)])
def test_dp_candidate_fates_match_the_oracle_attempt_log(local, cap, synth_index, synth_genome, monkeypatch):
    """bt2g_dp_extend's per-candidate fates: FAILED / SUCCEEDED exactly at the candidates the reference would start a backtrace
    from (= consume an RNG reseed), in order: what the exact policy needs from the DP kernel beyond the alignments."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bowtie2_b200 import Bt2Gpu, policy, synth
    from bowtie2_b200.lib import DP_PROBLEM, ReadBatch
    from oracle_lib import Oracle, oracle_dp
    # cap "1": the move-code kernels (sequential candidate loop); None: the default split H-byte kernels, whose tail screens
    # the candidates after the first alignment in parallel (end-to-end candidates all start in the last row, which lies
    # inside the gap barrier: no walk can mark another candidate's start cell, so the screening verdicts ARE the fates)
    g = Bt2Gpu(0)
    if cap is not None:
        g.set_dp_mode(int(cap))
    g.load_index_files(synth_index)
    g.set_scoring(local=local)
    O = Oracle(synth_index)
    sc = policy.Scoring.default(local)
    reads, quals, truth = synth.make_reads(synth_genome, 300, 100, seed=91, sub_rate=0.03, indel_rate=0.01)
    probs = np.zeros(len(reads), dtype=DP_PROBLEM)
    meta = []
    for i, r in enumerate(reads):
        c, pos, fw = int(truth[i][0]), int(truth[i][1]), int(truth[i][2]) > 0
        if c < 0:                                           # a random read: frame it anywhere
            c, pos, fw = 0, 1000 + i, True
        rdlen = len(r)
        minsc = sc.min_score(rdlen)
        tlen = len(synth_genome[c])
        found, rect = policy.frame_seed_extension_rect(pos, rdlen, tlen, sc.max_read_gaps(minsc, rdlen), sc.max_ref_gaps(minsc, rdlen),
                                                       sc.n_ceil(rdlen))
        probs[i] = (i, int(fw), c, rect.refl, rect.refr, rect.triml, rect.corel, rect.corer, minsc, sc.n_ceil_raw(rdlen), 0)
        meta.append(rect)
    summ, cands, alns, ops = g.dp_extend(ReadBatch.from_list(reads, quals), probs, max_cands=16384 if local else 512, max_alns=32)
    n_att = 0
    for i, r in enumerate(reads):
        d = oracle_dp(O, local, r, quals[i], bool(probs[i]["fw"]), int(probs[i]["tidx"]), meta[i], int(probs[i]["minsc"]),
                      int(probs[i]["nceil"]), max_cands=65536, max_alns=64, max_edits=16384, attempts=True)
        if not d["found"]:
            continue
        got = [(ci, int(cands[i][ci]["fate"])) for ci in range(int(summ[i]["ncand"])) if int(cands[i][ci]["fate"]) in (2, 3)]
        want = [(ci, 3 if ai >= 0 else 2) for (s, ai), ci in zip(d["attempts"], d["attempt_cands"])]
        assert got == want, (i, got[:6], want[:6])
        n_att += len(want)
    assert n_att > 100
    g.close()
