"""Host-side index reader (bt2g_index_file_*, no GPU): every array and scalar against an independent numpy parse of
the files bowtie2-build wrote (field order: SURVEY.md Appendix A / bt2_io.cpp:131-616), the --offrate override, the
endian switch, and the SAM header built from the stored names and lengths against the reference program's."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from bowtie2_b200.lib import IndexFile, load_library, sam_header
from conftest import GOLDEN
from oracle_lib import ref_bin


def _parse(base, ext):
    """independent parse of <base>.1/.2/.3/.4/.rev.1 (little-endian files)"""
    off = np.dtype("<u4") if ext == "bt2" else np.dtype("<u8")
    osz = off.itemsize
    out = {}

    def one(path, tag, want_names):
        b = open(path, "rb").read()
        pos = 0

        def i32():
            nonlocal pos
            v = int(np.frombuffer(b, "<i4", 1, pos)[0]); pos += 4
            return v

        def offv(n=1):
            nonlocal pos
            v = np.frombuffer(b, off, n, pos).copy(); pos += n * osz
            return v
        assert i32() == 1
        ln = int(offv()[0]); line_rate = i32(); i32(); off_rate = i32(); ftab_chars = i32(); i32()
        n_pat = int(offv()[0]); plen = offv(n_pat); n_frag = int(offv()[0]); rstarts = offv(3 * n_frag)
        side = 1 << line_rate
        nsides = ((ln // 4 + 1) + (side - 4 * osz) - 1) // (side - 4 * osz)
        ebwt = np.frombuffer(b, np.uint8, nsides * side, pos).copy(); pos += nsides * side
        zoff = int(offv()[0]); fchr = offv(5); ftab = offv((1 << (2 * ftab_chars)) + 1); eftab = offv(2 * ftab_chars)
        out.update({f"ebwt_{tag}": ebwt, f"ftab_{tag}": ftab, f"eftab_{tag}": eftab, f"z_off_{tag}": zoff})
        if want_names:
            names = b[pos:].split(b"\0", 1)[0].decode().split("\n")
            out.update(len=ln, line_rate=line_rate, off_rate=off_rate, ftab_chars=ftab_chars, n_pat=n_pat, n_frag=n_frag,
                       plen=plen, rstarts=rstarts, fchr=[int(x) for x in fchr], names=[x for x in names if x])
    one(f"{base}.1.{ext}", "fw", True)
    one(f"{base}.rev.1.{ext}", "bw", False)
    b = open(f"{base}.2.{ext}", "rb").read()
    out["offs"] = np.frombuffer(b, off, (len(b) - 4) // osz, 4).copy()
    b = open(f"{base}.3.{ext}", "rb").read()
    nrec = int(np.frombuffer(b, off, 1, 4)[0])
    rec = np.frombuffer(b, np.dtype([("off", off), ("len", off), ("first", "u1")]), nrec, 4 + osz)
    out.update(n_recs=nrec, rec_off=rec["off"].copy(), rec_len=rec["len"].copy(), rec_first=rec["first"].copy())
    out["ref_buf"] = np.frombuffer(open(f"{base}.4.{ext}", "rb").read(), np.uint8).copy()
    return out


def _check_equal(f, want, skip=()):
    sc = f.scalars()
    for k in ("len", "line_rate", "off_rate", "ftab_chars", "n_pat", "n_frag", "z_off_fw", "z_off_bw", "n_recs", "fchr"):
        if k not in skip:
            assert sc[k] == want[k], k
    for k in IndexFile._ARRAYS:
        if k in skip:
            continue
        a = f.array(k)
        w = want[k]
        if k == "rec_first":
            w = (w != 0).astype(np.uint8)
        assert a is not None and a.dtype.itemsize == w.dtype.itemsize and np.array_equal(a, w), k
    assert f.ref_names == want["names"]
    assert f.ref_lens == [int(x) for x in want["plen"]]


@pytest.mark.parametrize("which", ["lambda_index", "synth_index", "synth_index_large"])
def test_index_file_matches_independent_parse(which, request):
    base = request.getfixturevalue(which)
    ext = "bt2l" if which.endswith("large") else "bt2"
    want = _parse(base, ext)
    f = IndexFile(base)
    assert f.scalars()["off_size"] == (8 if ext == "bt2l" else 4)
    _check_equal(f, want)
    f.close()


def test_offrate_override(synth_index):
    want = _parse(synth_index, "bt2")
    r0 = want["off_rate"]
    f = IndexFile(synth_index, offrate=r0 + 3)
    assert f.scalars()["off_rate"] == r0 + 3
    assert np.array_equal(f.array("offs"), want["offs"][::8])
    _check_equal(f, want, skip=("off_rate", "offs"))
    # an override that is not sparser than the stored sample is ignored (bt2_io.cpp:222-224)
    g = IndexFile(synth_index, offrate=r0 - 1)
    _check_equal(g, want)
    g = IndexFile(synth_index, offrate=r0)
    _check_equal(g, want)


def _swap_index(src, dst, ext):
    """byte-swap the fields the reference's reader swaps (everything but ebwt[], the names and the .4 bytes)"""
    osz = 4 if ext == "bt2" else 8
    o_le, o_be = (np.dtype("<u4"), np.dtype(">u4")) if osz == 4 else (np.dtype("<u8"), np.dtype(">u8"))

    def sw(b, dt_le, dt_be, n, pos):
        return np.frombuffer(b, dt_le, n, pos).astype(dt_be).tobytes()
    for part in ("1", "rev.1"):
        b = open(f"{src}.{part}.{ext}", "rb").read()
        out = bytearray()
        pos = 0

        def take_i32(n=1):
            nonlocal pos
            out.extend(sw(b, "<i4", ">i4", n, pos)); v = np.frombuffer(b, "<i4", n, pos).copy(); pos += 4 * n
            return v

        def take_off(n=1):
            nonlocal pos
            out.extend(sw(b, o_le, o_be, n, pos)); v = np.frombuffer(b, o_le, n, pos).copy(); pos += osz * n
            return v
        take_i32()
        ln = int(take_off()[0]); hdr = take_i32(5); line_rate, ftab_chars = int(hdr[0]), int(hdr[3])
        n_pat = int(take_off()[0]); take_off(n_pat); n_frag = int(take_off()[0]); take_off(3 * n_frag)
        side = 1 << line_rate
        nsides = ((ln // 4 + 1) + (side - 4 * osz) - 1) // (side - 4 * osz)
        out.extend(b[pos:pos + nsides * side]); pos += nsides * side
        take_off(1); take_off(5); take_off((1 << (2 * ftab_chars)) + 1); take_off(2 * ftab_chars)
        out.extend(b[pos:])
        open(f"{dst}.{part}.{ext}", "wb").write(bytes(out))
    b = open(f"{src}.2.{ext}", "rb").read()
    open(f"{dst}.2.{ext}", "wb").write(sw(b, "<i4", ">i4", 1, 0) + sw(b, o_le, o_be, (len(b) - 4) // osz, 4))
    b = open(f"{src}.3.{ext}", "rb").read()
    nrec = int(np.frombuffer(b, o_le, 1, 4)[0])
    out = bytearray(sw(b, "<i4", ">i4", 1, 0) + sw(b, o_le, o_be, 1, 4))
    pos = 4 + osz
    for _ in range(nrec):
        out.extend(sw(b, o_le, o_be, 2, pos)); out.append(b[pos + 2 * osz]); pos += 2 * osz + 1
    open(f"{dst}.3.{ext}", "wb").write(bytes(out))
    shutil.copy(f"{src}.4.{ext}", f"{dst}.4.{ext}")


@pytest.mark.parametrize("which", ["synth_index", "synth_index_large"])
def test_endian_switched_files_read_the_same(which, request, tmp_path):
    base = request.getfixturevalue(which)
    ext = "bt2l" if which.endswith("large") else "bt2"
    dst = str(tmp_path / "swapped")
    _swap_index(base, dst, ext)
    _check_equal(IndexFile(dst), _parse(base, ext))
    # and the override applies to switched files as well
    f = IndexFile(dst, offrate=_parse(base, ext)["off_rate"] + 1)
    assert np.array_equal(f.array("offs"), _parse(base, ext)["offs"][::2])


def test_errors(tmp_path, synth_index):
    with pytest.raises(RuntimeError):
        IndexFile(str(tmp_path / "nothing"))
    # truncated .2 file
    for part in ("1", "2", "3", "4", "rev.1"):
        shutil.copy(f"{synth_index}.{part}.bt2", str(tmp_path / f"t.{part}.bt2"))
    b = open(str(tmp_path / "t.2.bt2"), "rb").read()
    open(str(tmp_path / "t.2.bt2"), "wb").write(b[:len(b) // 2])
    with pytest.raises(RuntimeError, match="short offs"):
        IndexFile(str(tmp_path / "t"))
    open(str(tmp_path / "t.2.bt2"), "wb").write(b"\x02\x00\x00\x00" + b[4:])
    with pytest.raises(RuntimeError, match="sentinel"):
        IndexFile(str(tmp_path / "t"))


@pytest.mark.parametrize("which,sam", [("lambda_index", "lambda_U_sensitive.sam"), (None, "rep_P_sensitive.sam")])
def test_sam_header_matches_golden(which, sam, request, tmp_path):
    if which is None:
        base = str(tmp_path / "rep")
        exe = ref_bin("bowtie2-build-s")
        if not os.path.exists(exe):
            pytest.skip("reference builder not built")
        subprocess.check_call([exe, "--seed", "0", "--quiet", os.path.join(GOLDEN, "rep_genome.fa"), base])
    else:
        base = request.getfixturevalue(which)
    f = IndexFile(base)
    want = "".join(l for l in open(os.path.join(GOLDEN, sam)) if l.startswith("@"))       # @PG was dropped from the fixtures
    lib = load_library()
    assert sam_header(lib, f.ref_names, f.ref_lens) == want
    assert sam_header(lib, f.ref_names, f.ref_lens, pg_cl="bowtie2-align-s -x y").endswith(
        '@PG\tID:bowtie2\tPN:bowtie2\tVN:2.5.5\tCL:"bowtie2-align-s -x y"\n')


def test_sam_header_with_n_gaps_matches_reference_program(synth_index, tmp_path):
    """contigs with N gaps: @SQ LN is plen[] as stored in the index; the reference run on an empty read file prints the header only"""
    exe = ref_bin("bowtie2-align-s")
    if not os.path.exists(exe):
        pytest.skip("reference aligner not built")
    fq = tmp_path / "empty.fq"
    fq.write_text("")
    out = subprocess.run([exe, "-x", synth_index, "-U", str(fq)], capture_output=True, text=True, check=True).stdout
    want = "".join(l + "\n" for l in out.split("\n") if l.startswith("@") and not l.startswith("@PG"))
    f = IndexFile(synth_index)
    assert sam_header(load_library(), f.ref_names, f.ref_lens) == want
