"""csrc/xengine.cuh on the CPU: the fixed-memory state machine of the exact search policy (the code k_xe_step runs on the device)
driven by csrc/xengine_host.cpp over the entry-point table answered by the CPU oracle (tests/fake_gpu.py).  Its SAM must equal the
reference program's golden files; units that outgrow a capacity fall back to the coroutine engine (same records)."""
import os

import pytest

from bowtie2_b200.lib import ReadBatch, load_library, policy_align, policy_params, sam_format
from conftest import GOLDEN, read_fastq_codes
from test_policy_engine_cpp import _table

ENTRY = "bt2g_xengine_align_host"


@pytest.mark.parametrize("fixture,index,ref_names,local", [
    ("lambda_U_sensitive", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], False),
    ("lambda_U_local", "lambda_index", ["gi|9626243|ref|NC_001416.1|"], True),
    ("rep_U_sensitive", "rep_index", ["ctg1", "ctg2"], False),
])
def test_state_machine_unpaired_sam_identical_to_golden(fixture, index, ref_names, local, request):
    base = request.getfixturevalue(index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, fixture + ".sam")) if not l.startswith("@")]
    names, reads, quals = read_fastq_codes(os.path.join(GOLDEN, fixture.split("_")[0] + "_reads_1.fq"), len(golden))
    be, keep, fake = _table(base, local)
    lib = load_library()
    batch = ReadBatch.from_list(reads, quals)
    res, ops, _, (units, fallbacks, requests) = policy_align(lib, be, policy_params("sensitive", local=local), batch, names, entry=ENTRY)
    lines = sam_format(lib, batch, res, ops, ref_names, read_names=names, local=local).rstrip("\n").split("\n")
    assert lines == golden
    assert units == len(reads) and fallbacks * (4 if local else 20) <= units, (units, fallbacks)


@pytest.mark.parametrize("fixture,index,ref_names", [
    ("lambda", "lambda_index", ["gi|9626243|ref|NC_001416.1|"]),
    ("rep", "rep_index", ["ctg1", "ctg2"]),
])
def test_state_machine_paired_sam_identical_to_golden(fixture, index, ref_names, request):
    base = request.getfixturevalue(index)
    golden = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, f"{fixture}_P_sensitive.sam")) if not l.startswith("@")]
    n = len(golden) // 2
    n1, r1, q1 = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_1.fq"), n)
    n2, r2, q2 = read_fastq_codes(os.path.join(GOLDEN, f"{fixture}_reads_2.fq"), n)
    il = lambda a, b: [x for p in zip(a, b) for x in p]
    R, Q, N = il(r1, r2), il(q1, q2), il(n1, n2)
    be, keep, fake = _table(base)
    lib = load_library()
    batch = ReadBatch.from_list(R, Q)
    res, ops, pairs, (units, fallbacks, requests) = policy_align(lib, be, policy_params("sensitive", paired=True), batch, N, entry=ENTRY)
    lines = sam_format(lib, batch, res, ops, ref_names, read_names=N, pairs=pairs).rstrip("\n").split("\n")
    assert lines == golden
    assert fallbacks * 20 <= units, (units, fallbacks)
