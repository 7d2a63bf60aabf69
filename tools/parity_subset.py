#!/usr/bin/env python
"""Parity-subset run of the compiled exact-policy engine ("a 100k-read prefix ... full SAM diff").
  python tools/parity_subset.py U 100000 sensitive 100         # configs[1] parameters
  python tools/parity_subset.py P 50000 very-sensitive 150     # configs[2]: 2x150, --very-sensitive
  python tools/parity_subset.py U 20000 very-sensitive 300 local
Builds a 5 Mbp four-contig genome with repeats and N gaps, runs the reference PROGRAM (oracle/_ref) and bt2g_policy_align over
the oracle-backed entry-point table (oracle/bt2_oracle_table.c; TABLE=python: tests/fake_gpu.py), and diffs every SAM record.  LARGE=1: a .bt2l index (bowtie2-build-l / bowtie2-align-l), configuration 5's format.  CPU only; minutes, not part of pytest."""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle_lib import Oracle, ref_bin
from fake_gpu import FakeGpu, backend_table
from oracle_lib import oracle_policy_table
from bowtie2_b200 import synth
from bowtie2_b200.lib import ReadBatch, load_library, sam_format, policy_align, policy_params
paired = sys.argv[1] == 'P'; N = int(sys.argv[2]); preset = sys.argv[3]; rdlen = int(sys.argv[4]); LOCAL = len(sys.argv) > 5 and sys.argv[5] == 'local'
tmp = tempfile.mkdtemp()
genome = synth.make_genome(n_contigs=4, contig_len=1250000, seed=20260922, repeat_frac=0.03, repeat_len=2000, repeat_copies=60, n_gap=1000)
fa = os.path.join(tmp, 'g.fa'); synth.write_fasta(fa, genome); base = os.path.join(tmp, 'g')
t0 = time.time(); subprocess.check_call([ref_bin('bowtie2-build-l' if os.environ.get('LARGE') else 'bowtie2-build-s'), '--seed', '0', '--quiet', '--threads', '8', fa, base]); print('index built', time.time() - t0)
if paired:
    reads, quals, _ = synth.make_pairs(genome, N, rdlen, seed=5, sub_rate=0.005, indel_rate=0.0005, ins_mean=350, ins_sd=30)
    names = [f"r{i // 2}" for i in range(2 * N)]
    f1, f2 = os.path.join(tmp, 'r1.fq'), os.path.join(tmp, 'r2.fq')
    synth.write_fastq(f1, reads[0::2], quals[0::2]); synth.write_fastq(f2, reads[1::2], quals[1::2])
    io = ['-1', f1, '-2', f2]
else:
    reads, quals, _ = synth.make_reads(genome, N, rdlen, seed=5, sub_rate=0.005, indel_rate=0.0005)
    names = [f"r{i}" for i in range(N)]
    fq = os.path.join(tmp, 'r.fq'); synth.write_fastq(fq, reads, quals); io = ['-U', fq]
t0 = time.time()
out = subprocess.check_output([ref_bin('bowtie2-align-l' if os.environ.get('LARGE') else 'bowtie2-align-s'), *(['--local'] if LOCAL else []), '--' + preset, '--seed', '0', '-p', '8', '--reorder', '-x', base] + io, stderr=subprocess.DEVNULL).decode()
print('reference program: %.1f s on 8 threads' % (time.time() - t0))
want = [l for l in out.split('\n') if l and not l.startswith('@')]
lib = load_library()
if os.environ.get('TABLE') == 'python':      # the Python stand-in device (tests/fake_gpu.py)
    fake = FakeGpu(Oracle(base)); fake.set_scoring(LOCAL); be, keep = backend_table(fake)
else:                                         # the C table over the oracle (oracle/bt2_oracle_table.c)
    be, keep = oracle_policy_table(Oracle(base), LOCAL, 8 if os.environ.get('LARGE') else 4)
batch = ReadBatch.from_list(reads, quals)
t0 = time.time()
ENTRY = os.environ.get('ENTRY', 'bt2g_policy_align')      # ENTRY=bt2g_xengine_align_host: the state machine of csrc/xengine.cuh
res, ops, pairs, stats = policy_align(lib, be, policy_params(preset, local=LOCAL, paired=paired, host_threads=8), batch, names, entry=ENTRY)
print(ENTRY, 'over the oracle-backed table: %.1f s' % (time.time() - t0), 'stats', stats)
lines = sam_format(lib, batch, res, ops, [f"chr{k+1}" for k in range(4)], read_names=names, pairs=pairs, threads=8, local=LOCAL).rstrip('\n').split('\n')
bad = [i for i in range(len(want)) if lines[i] != want[i]]
print('records identical: %d of %d' % (len(want) - len(bad), len(want)))
for i in bad[:3]:
    g, w = lines[i].split('\t'), want[i].split('\t'); print('  got ', g[:9], g[11:]); print('  want', w[:9], w[11:])
