"""SwDriver::extend: C restatement vs reference (CPU) and CUDA vs restatement (GPU)."""
import numpy as np
import pytest

from bowtie2_b200 import synth
from oracle_lib import Oracle, Reference, extend_both, have_reference


def _cases(genome, n=150, rdlen=90, L=20, ival=7):
    reads, _, _ = synth.make_reads(genome, n, rdlen, seed=21, sub_rate=0.03, indel_rate=0.003)
    rng = np.random.default_rng(4)
    for r in reads[:20]:
        r[rng.integers(0, len(r))] = 4
    return reads, L, ival


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_extend_oracle_vs_reference(synth_index, synth_genome):
    O, R = Oracle(synth_index), Reference(synth_index)
    reads, L, ival = _cases(synth_genome)
    nchk = 0
    for r in reads:
        n, out = O.seed_search(r, L, ival, 0, 32)
        for strand in range(2):
            for k in range(n):
                rg = out[strand, k]
                if rg[1] <= rg[0]:
                    continue
                a = extend_both(O, r, strand == 0, k * ival, L, rg)
                b = extend_both(R, r, strand == 0, k * ival, L, rg)
                assert a == b, (strand, k, a, b)
                nchk += 1
    assert nchk > 500


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("which", ["small", "large"])
@pytest.mark.parametrize("through_text", [True, False])
def test_extend_gpu_vs_oracle(which, through_text, gpu, synth_index, synth_index_large, synth_genome):
    """both extension paths of the library (unique hits compared with the packed reference = default; index walk) against the
    oracle; incl. reads cut from the two ends of the joined text (the "$" row) and reads ending in Ns there"""
    from bowtie2_b200.lib import ReadBatch
    base = synth_index if which == "small" else synth_index_large
    gpu.load_index_files(base)
    gpu.set_extend_mode(through_text)
    O = Oracle(base)
    reads, L, ival = _cases(synth_genome)
    first = np.array([c for c in synth_genome[0] if c < 4][:90], dtype=np.uint8)
    last = np.array([c for c in synth_genome[-1] if c < 4][-90:], dtype=np.uint8)
    for e in (first, last):
        reads.append(e.copy())
        for cut in (30, 55):
            x = e.copy(); x[:cut] = 4; reads.append(x)          # Ns run over the text's start
            y = e.copy(); y[-cut:] = 4; reads.append(y)         # ... and over its end
        reads.append((3 - e[::-1]).astype(np.uint8))            # reverse complement
    batch = ReadBatch.from_list(reads)
    ranges, ns = gpu.seed_search(batch, L, ival, 0, 32)
    ext = gpu.extend_exact(batch, L, ival, 0, 32, ranges)
    nchk = 0
    for i, r in enumerate(reads):
        for strand in range(2):
            for k in range(int(ns[i])):
                rg = ranges[i, strand, k]
                if rg[1] <= rg[0]:
                    assert tuple(ext[i, strand, k]) == (0, 0)
                    continue
                want = extend_both(O, r, strand == 0, k * ival, L, rg)
                assert tuple(int(x) for x in ext[i, strand, k]) == want, (i, strand, k)
                nchk += 1
    assert nchk > 500
    gpu.set_extend_mode(True)
