"""The reference's own regression corpus (scripts/test/simple_tests.pl, 272 cases: tiny references, hand-made reads with
gaps at the ends, Ns, overlapping / containing / dovetailing mates, repeats ...) as inputs: every case whose options the
policy engine models is run through the reference PROGRAM and through the engine (oracle backend) and the two SAM files
must be identical, record for record.  Cases with options outside the engine's scope (raw --policy strings, read
trimming, --overhang, -N 1, non-FASTQ input formats) are counted and skipped."""
import json
import os
import re
import shlex
import subprocess

import numpy as np
import pytest

from bowtie2_b200 import policy
from bowtie2_b200.policy_engine import PairedPolicyEngine, PolicyEngine
from conftest import GOLDEN
from oracle_lib import Oracle, have_reference, ref_bin
from policy_backend import OracleBackend
from test_policy_engine import _multi_sam, _multi_sam_pairs

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
FUNC = {"C": policy.SIMPLE_FUNC_CONST, "L": policy.SIMPLE_FUNC_LINEAR, "S": policy.SIMPLE_FUNC_SQRT, "G": policy.SIMPLE_FUNC_LOG}
FORMATS = ("fastq", "tabbed", "fasta", "qseq", "raw", "cline_reads", "cont_fasta_reads", "fastq1", "fasta1", "raw1", "qseq1",
           "cline_reads1", "should_abort", "lines")


def _func(text, dflt_min=0.0, dflt_max=float("inf")):
    t = text.split(",")
    return policy.SimpleFunc(FUNC[t[0]], float(t[1]), float(t[2]) if len(t) > 2 else 0.0)


class _BwaSwLike:
    """--bwa-sw-like's threshold: (TAlScore)max<float>(a*T, a*c*log(len)) with a = 1, T = 30, c = 5.5f"""
    def f(self, x):
        import math
        f32 = policy._F32
        return float(int(max(f32(30.0), f32(5.5 * math.log(x)))))

    def fi(self, x):
        return int(self.f(x))


def _options(case):
    """-> (reference arguments, engine kwargs, scoring, pe policy kwargs) or None when an option is out of scope"""
    toks = shlex.split((case.get("args") or "") + " " + (case.get("report") if case.get("report") is not None else "-a"))
    local = "--local" in toks or "--bwa-sw-like" in toks
    sc = policy.Scoring.default(local)
    kw, pe = {}, {}
    i = 0
    while i < len(toks):
        t = toks[i]
        val = None
        if "=" in t and t.startswith("--"):
            t, val = t.split("=", 1)

        def arg():
            nonlocal i
            if val is not None:
                return val
            i += 1
            return toks[i]
        if t in ("--local", "--quiet"):
            pass
        elif t == "-a":
            kw["all_hits"] = True
        elif t == "-k":
            kw["k"] = int(arg())
        elif t == "-M":
            kw["mhits"] = int(arg())
        elif t == "-L":
            kw["seed_len"] = int(arg())
        elif t == "-i":
            kw["ival"] = _func(arg())
        elif t == "-D":
            kw["dp_fail_streak"] = int(arg())
        elif t == "-R":
            kw["seed_rounds"] = int(arg())
        elif t == "--seed":
            kw["seed"] = int(arg())
        elif t == "--multiseed":
            f = arg().split(",")                       # mms, length, interval function...
            if int(f[0]) != 0:
                return None
            kw["seed_len"] = int(f[1])
            if len(f) >= 4:
                kw["ival"] = policy.SimpleFunc(FUNC[f[2]], float(f[3]), float(f[4]) if len(f) > 4 else 0.0)
        elif t in ("--score-min", "--min-score"):
            sc.score_min_func = _func(arg())
        elif t == "--bwa-sw-like":
            # bt2_search.cpp:1114-1126 (scoring MA=1;MMP=C3;RDG=5,2;RFG=5,2, local) and :3341-3350 (threshold a*max{T, c*log(l)})
            sc.match_bonus, sc.mmp_max, sc.mmp_min = 1, 3, 3
            sc.rdgap_const, sc.rdgap_linear, sc.rfgap_const, sc.rfgap_linear = 5, 2, 5, 2
            sc.score_min_func = _BwaSwLike()
        elif t == "--ignore-quals":
            sc.mmp_min = sc.mmp_max
        elif t == "--mp":
            f = arg().split(",")
            sc.mmp_max = int(f[0])
            if len(f) > 1:
                sc.mmp_min = int(f[1])
        elif t == "--nofw":
            kw["nofw"] = True
        elif t == "--norc":
            kw["norc"] = True
        elif t == "-I":
            pe["minfrag"] = int(arg())
        elif t == "-X":
            pe["maxfrag"] = int(arg())
        elif t in ("--ff", "--fr", "--rf"):
            pe["pol"] = {"--ff": 1, "--fr": 3, "--rf": 4}[t]
        elif t == "--no-mixed":
            kw["mixed"] = False
        elif t == "--no-discordant":
            kw["discord"] = False
        elif t == "--no-contain":
            pe["contain_ok"] = False
        elif t == "--no-overlap":
            pe["olap_ok"] = False
        elif t == "--dovetail":
            pe["dovetail_ok"] = True
        elif t == "--no-dovetail":
            pe["dovetail_ok"] = False
        elif t in ("-3", "-5", "--trim3", "--trim5"):
            kw["_trim3" if t in ("-3", "--trim3") else "_trim5"] = int(arg())
        elif t == "--trim-to":
            v = arg()                                  # [3:|5:]N: trim the 3' (default) or 5' end so that N bases remain
            end, num = ("3", v) if ":" not in v else tuple(v.split(":"))
            kw["_trimto"] = (end, int(num))
        elif t in ("-u", "-s"):
            kw["_upto" if t == "-u" else "_skip"] = int(arg())
        elif t == "--rdg":
            f = arg().split(",")
            sc.rdgap_const = int(f[0])
            if len(f) > 1:
                sc.rdgap_linear = int(f[1])
        elif t == "--rfg":
            f = arg().split(",")
            sc.rfgap_const = int(f[0])
            if len(f) > 1:
                sc.rfgap_linear = int(f[1])
        elif t == "--np":
            sc.n_pen = int(arg())
        elif t == "--ma":
            sc.match_bonus = int(arg())
        elif t == "--n-ceil":
            sc.n_ceil_over = _func(arg())
        elif t == "--policy":
            # SeedAlignmentPolicy::parseString (aligner_seed_policy.cpp): integer-valued settings only
            for kv in arg().replace("\\;", ";").replace("\\", "").split(";"):
                if not kv:
                    continue
                key, v = kv.split("=", 1)
                try:
                    if key == "SEED":
                        if int(v.split(",")[0]) != 0:
                            return None
                        if "," in v:
                            kw["seed_len"] = int(v.split(",")[1])
                    elif key == "SEEDLEN":
                        kw["seed_len"] = int(v)
                    elif key == "IVAL":
                        kw["ival"] = _func(v)
                    elif key == "MIN":
                        sc.score_min_func = _func(v)
                    elif key == "NCEIL":
                        sc.n_ceil_over = _func(v)
                    elif key == "MMP":
                        if v.startswith("C"):
                            sc.mmp_max = sc.mmp_min = int(v[1:])
                        elif v != "Q":
                            return None
                    elif key == "NP":
                        if not v.startswith("C"):
                            return None
                        sc.n_pen = int(v[1:])
                    elif key in ("RDG", "RFG"):
                        f = v.split(",")
                        a, b = int(f[0]), (int(f[1]) if len(f) > 1 else None)
                        if key == "RDG":
                            sc.rdgap_const = a
                            sc.rdgap_linear = b if b is not None else sc.rdgap_linear
                        else:
                            sc.rfgap_const = a
                            sc.rfgap_linear = b if b is not None else sc.rfgap_linear
                    else:
                        return None
                except ValueError:
                    return None
        else:
            return None
        i += 1
    if "all_hits" in kw and "k" in kw:
        del kw["all_hits"]
    if ("all_hits" in kw or "k" in kw) and "mhits" in kw:
        del kw["mhits"]                                 # -k / -a switch -M off (bt2_search.cpp:1772-1774)
    return toks, kw, sc, pe, local


def _usable(case):
    if any(k in case for k in FORMATS) or not ("reads" in case or "mate1s" in case):
        return False
    seqs = (case.get("reads") or []) + (case.get("mate1s") or []) + (case.get("mate2s") or [])
    if any(re.search(r"[^ACGTN]", s) for s in seqs):
        return False
    return _options(case) is not None


_INDEX_CACHE = {}


def _index_for(tmp_path, refs):
    """bowtie2-build once per distinct reference set"""
    key = tuple(refs)
    if key not in _INDEX_CACHE:
        d = tmp_path / f"ix{len(_INDEX_CACHE)}"
        d.mkdir()
        fa, base = str(d / "ref.fa"), str(d / "ref")
        with open(fa, "w") as f:
            for k, r in enumerate(refs):
                f.write(f">{k}\n{r}\n")
        subprocess.check_call([ref_bin("bowtie2-build-s"), "--quiet", fa, base])
        _INDEX_CACHE[key] = base
    return _INDEX_CACHE[key]


def _cases():
    return json.load(open(os.path.join(GOLDEN, "simple_tests_cases.json")))


def _codes(s):
    return np.array([CODE.get(c, 4) for c in s], dtype=np.uint8)


def _write_fq(path, seqs, quals, names, suffix=""):
    with open(path, "w") as f:
        for s, q, n in zip(seqs, quals, names):
            f.write(f"@{n}{suffix}\n{s}\n+\n{q}\n")


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_reference_regression_corpus(tmp_path):
    cases = _cases()
    assert len(cases) == 272
    n_run = n_pairs = n_reads = 0
    skipped = 0
    for ci, case in enumerate(cases):
        if not _usable(case):
            skipped += 1
            continue
        toks, kw, sc, pe_kw, local = _options(case)
        skip, upto = kw.pop("_skip", 0), kw.pop("_upto", None)
        t5, t3 = kw.pop("_trim5", 0), kw.pop("_trim3", 0)
        tt = kw.pop("_trimto", None)

        def trim(x):                                       # -5 / -3 / --trim-to: bases removed before alignment (and from SEQ / QUAL)
            if tt is not None:
                return x if len(x) <= tt[1] else (x[:tt[1]] if tt[0] == "3" else x[len(x) - tt[1]:])
            return x[t5:len(x) - t3] if t3 else x[t5:]
        d = tmp_path / f"c{ci}"
        d.mkdir()
        base = _index_for(tmp_path, case["ref"])
        ref_names = [str(k) for k in range(len(case["ref"]))]
        paired = "mate1s" in case
        if paired and ("mate1fw" in case or "mate2fw" in case) and "pol" not in pe_kw:
            m1, m2 = case.get("mate1fw", 1), case.get("mate2fw", 0)
            if m1 == m2:
                toks, pe_kw["pol"] = toks + ["--ff"], 1
            elif not m1:
                toks, pe_kw["pol"] = toks + ["--rf"], 4
        if case.get("norc"):
            pass                                         # harness-side: do not also test the reverse complements
        cmd = [ref_bin("bowtie2-align-s"), "--quiet", "-p", "1", "-x", base] + [t for t in toks if t != "--quiet"]
        if "--seed" not in " ".join(toks):
            cmd += ["--seed", "0"]
        kw.setdefault("seed", 0)
        backend = OracleBackend(Oracle(base), local=local, scoring=sc)
        if not paired:
            seqs = case["reads"]
            quals = [(case.get("quals") or [None] * len(seqs))[k] or "I" * len(s) for k, s in enumerate(seqs)]
            names = [(case.get("names") or [None] * len(seqs))[k] or f"r{k}" for k in range(len(seqs))]
            fq = str(d / "r.fq")
            _write_fq(fq, seqs, quals, names)
            out = subprocess.run(cmd + ["-U", fq], capture_output=True, text=True)
            if out.returncode != 0:
                skipped += 1
                continue
            want = [l for l in out.stdout.split("\n") if l and not l.startswith("@")]
            eng = PolicyEngine(backend, "sensitive", sc=sc, local=local, **{k: v for k, v in kw.items() if k not in ("mixed", "discord")})
            R = [trim(_codes(s)) for s in seqs]
            Q = [trim(np.frombuffer(q.encode(), dtype=np.uint8)) for q in quals]
            sel = range(len(R))[skip:][:upto] if upto is not None else range(len(R))[skip:]     # -s / -u
            R, Q, names = [R[k] for k in sel], [Q[k] for k in sel], [names[k] for k in sel]
            outs = [eng.align_read(R[k], Q[k], names[k]) for k in range(len(R))]
            lines = _multi_sam(outs, R, Q, names, ref_names, local=local)
            n_reads += len(R)
        else:
            s1, s2 = case["mate1s"], case["mate2s"]
            q1 = [(case.get("qual1s") or [None] * len(s1))[k] or "I" * len(s) for k, s in enumerate(s1)]
            q2 = [(case.get("qual2s") or [None] * len(s2))[k] or "I" * len(s) for k, s in enumerate(s2)]
            names = [(case.get("names") or [None] * len(s1))[k] or f"r{k}" for k in range(len(s1))]
            f1, f2 = str(d / "r1.fq"), str(d / "r2.fq")
            _write_fq(f1, s1, q1, names, "/1")
            _write_fq(f2, s2, q2, names, "/2")
            out = subprocess.run(cmd + ["-1", f1, "-2", f2], capture_output=True, text=True)
            if out.returncode != 0:
                skipped += 1
                continue
            want = [l for l in out.stdout.split("\n") if l and not l.startswith("@")]
            eng = PairedPolicyEngine(backend, "sensitive", sc=sc, local=local, pe=policy.PairedEndPolicy(local=local, **pe_kw), **kw)
            R = [trim(x) for p in zip((_codes(s) for s in s1), (_codes(s) for s in s2)) for x in p]
            Q = [trim(np.frombuffer(x.encode(), dtype=np.uint8)) for p in zip(q1, q2) for x in p]
            N = [x for k in range(len(s1)) for x in (names[k] + "/1", names[k] + "/2")]
            sel = range(len(s1))[skip:][:upto] if upto is not None else range(len(s1))[skip:]
            R = [R[2 * k + j] for k in sel for j in (0, 1)]
            Q = [Q[2 * k + j] for k in sel for j in (0, 1)]
            N = [N[2 * k + j] for k in sel for j in (0, 1)]
            outs = [eng.align_pair(R[2 * k], Q[2 * k], N[2 * k], R[2 * k + 1], Q[2 * k + 1], N[2 * k + 1]) for k in range(len(R) // 2)]
            lines = _multi_sam_pairs(outs, R, Q, N, ref_names, local=local)
            n_pairs += len(s1)
        assert lines == want, (ci, case.get("name"), case.get("args"), case.get("report"),
                               next(((a, b) for a, b in zip(lines, want) if a != b), (len(lines), len(want))))
        n_run += 1
    assert n_run >= 130, (n_run, skipped)
    print(f"simple_tests.pl corpus: {n_run} cases identical ({n_reads} reads, {n_pairs} pairs), {skipped} outside the engine's options")


FMT_KEYS = {"fasta": ("fasta", "-f"), "raw": ("raw", "-r"), "tabbed": ("tab5", "--tab5"), "qseq": ("qseq", "--qseq"),
            "cline_reads": ("cline", "-c"), "fastq": ("fastq", "-q")}


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_reference_corpus_read_formats(tmp_path):
    """the corpus' input-format cases (FASTA, raw, tab-delimited, qseq, command-line reads, FASTQ with odd line ends), unpaired
    and as mate files: bowtie2_b200.align.parse_reads / the FASTQ parser feed the engine, the reference reads the file itself;
    names, sequences, qualities and therefore the SAM must agree"""
    from bowtie2_b200.align import parse_reads
    from bowtie2_b200.lib import fastq_parse, load_library
    lib = load_library()
    n_run = skipped = 0
    why = []
    for ci, case in enumerate(_cases()):
        key = next((k for k in FMT_KEYS if k in case or k + "1" in case), None)
        if key is None or case.get("should_abort") or _options({k: v for k, v in case.items() if k in ("args", "report")}) is None:
            continue
        fmt, farg = FMT_KEYS[key]
        toks, kw, sc, pe_kw, local = _options(case)
        skip, upto = kw.pop("_skip", 0), kw.pop("_upto", None)
        t5, t3 = kw.pop("_trim5", 0), kw.pop("_trim3", 0)
        tt = kw.pop("_trimto", None)

        def trim(x, t5=t5, t3=t3, tt=tt):
            if tt is not None:
                return x if len(x) <= tt[1] else (x[:tt[1]] if tt[0] == "3" else x[len(x) - tt[1]:])
            return x[t5:len(x) - t3] if t3 else x[t5:]
        pick = lambda lst: (lst[skip:][:upto] if upto is not None else lst[skip:])
        paired_files = key + "1" in case
        d = tmp_path / f"f{ci}"
        d.mkdir()
        base = _index_for(tmp_path, case["ref"])
        ref_names = [str(k) for k in range(len(case["ref"]))]
        cmd = [ref_bin("bowtie2-align-s"), "--quiet", "-p", "1", "--seed", "0", "-x", base] + [t for t in toks if t != "--quiet"] + [farg]
        kw.setdefault("seed", 0)

        def load(text):
            if fmt == "fastq":
                b, names, used = fastq_parse(lib, text.encode() + (b"" if text.endswith("\n") else b"\n"))
                rd = [(b.seq[int(b.off[i]):int(b.off[i + 1])], b.qual[int(b.off[i]):int(b.off[i + 1])]) for i in range(b.n)]
                return [n.split()[0] if n.split() else n for n in names], [r[0] for r in rd], [r[1] for r in rd], None
            names, seqs, quals, m2 = parse_reads(text, fmt)
            R = [_codes(s.upper().replace(".", "N")) for s in seqs]
            Q = [np.frombuffer((q if q is not None else "I" * len(s)).encode(), dtype=np.uint8) for s, q in zip(seqs, quals)]
            return names, R, Q, m2
        backend = OracleBackend(Oracle(base), local=local, scoring=sc)
        try:
            if paired_files:
                t1, t2 = case[key + "1"], case[key + "2"]
                if fmt == "cline":
                    out = subprocess.run(cmd + ["-1", t1, "-2", t2], capture_output=True, text=True)
                else:
                    p1, p2 = str(d / "m1.txt"), str(d / "m2.txt")
                    open(p1, "w").write(t1); open(p2, "w").write(t2)
                    out = subprocess.run(cmd + ["-1", p1, "-2", p2], capture_output=True, text=True)
                n1, R1, Q1, _ = load(t1)
                n2, R2, Q2, _ = load(t2)
                n1, R1, Q1, n2, R2, Q2 = (pick(n1), pick([trim(x) for x in R1]), pick([trim(x) for x in Q1]),
                                          pick(n2), pick([trim(x) for x in R2]), pick([trim(x) for x in Q2]))
                strip = lambda n: n[:-2] if n.endswith(("/1", "/2")) else n
                N = [x for a, b in zip(n1, n2) for x in (strip(a), strip(b))]
                R = [x for p in zip(R1, R2) for x in p]
                Q = [x for p in zip(Q1, Q2) for x in p]
                eng = PairedPolicyEngine(backend, "sensitive", sc=sc, local=local, pe=policy.PairedEndPolicy(local=local, **pe_kw), **kw)
                outs = [eng.align_pair(R[2 * k], Q[2 * k], N[2 * k], R[2 * k + 1], Q[2 * k + 1], N[2 * k + 1]) for k in range(len(R1))]
                lines = _multi_sam_pairs(outs, R, Q, N, ref_names, local=local)
            else:
                text = case[key]
                if fmt == "cline":
                    out = subprocess.run(cmd + ["-U", text], capture_output=True, text=True)
                else:
                    p = str(d / "reads.txt")
                    open(p, "w").write(text)
                    # --tab5 / --tab6 take the file themselves (and may hold pairs); the other formats go through -U
                    out = subprocess.run(cmd + ([p] if fmt == "tab5" else ["-U", p]), capture_output=True, text=True)
                names, R, Q, m2 = load(text)
                if m2 is not None:
                    if any(x is None for x in m2):
                        skipped += 1                               # a file mixing pairs and single reads
                        why.append((ci, key, "mixed paired / unpaired records"))
                        continue
                    R2 = [_codes(x[1].upper().replace(".", "N")) for x in m2]
                    Q2 = [np.frombuffer(x[2].encode(), dtype=np.uint8) for x in m2]
                    N = [x for k in range(len(names)) for x in (names[k], m2[k][0])]
                    N = [x[:-2] if x.endswith(("/1", "/2")) else x for x in N]
                    Rp = [x for pp in zip(pick([trim(x) for x in R]), pick([trim(x) for x in R2])) for x in pp]
                    Qp = [x for pp in zip(pick([trim(x) for x in Q]), pick([trim(x) for x in Q2])) for x in pp]
                    Np = [x for k in (range(len(names))[skip:][:upto] if upto is not None else range(len(names))[skip:]) for x in N[2 * k:2 * k + 2]]
                    eng = PairedPolicyEngine(backend, "sensitive", sc=sc, local=local, pe=policy.PairedEndPolicy(local=local, **pe_kw), **kw)
                    outs = [eng.align_pair(Rp[2 * k], Qp[2 * k], Np[2 * k], Rp[2 * k + 1], Qp[2 * k + 1], Np[2 * k + 1]) for k in range(len(Rp) // 2)]
                    lines = _multi_sam_pairs(outs, Rp, Qp, Np, ref_names, local=local)
                    if out.returncode != 0:
                        skipped += 1
                        why.append((ci, key, "reference exit " + str(out.returncode)))
                        continue
                    want = [l for l in out.stdout.split("\n") if l and not l.startswith("@") and l.count("\t") >= 10]
                    assert lines == want, (ci, key, case.get("name"), next(((a, b) for a, b in zip(lines, want) if a != b), (len(lines), len(want))))
                    n_run += 1
                    continue
                names, R, Q = pick(names), pick([trim(x) for x in R]), pick([trim(x) for x in Q])
                eng = PolicyEngine(backend, "sensitive", sc=sc, local=local, **{k: v for k, v in kw.items() if k not in ("mixed", "discord")})
                outs = [eng.align_read(R[k], Q[k], names[k]) for k in range(len(R))]
                lines = _multi_sam(outs, R, Q, names, ref_names, local=local)
        except Exception as e:                                     # an input the readers do not model: count, do not hide
            skipped += 1
            why.append((ci, key, repr(e)[:80]))
            continue
        if out.returncode != 0:
            skipped += 1
            why.append((ci, key, "reference exit " + str(out.returncode)))
            continue
        want = [l for l in out.stdout.split("\n") if l and not l.startswith("@") and l.count("\t") >= 10]   # (@PG may span lines)
        assert lines == want, (ci, key, case.get("name"), next(((a, b) for a, b in zip(lines, want) if a != b), (len(lines), len(want))))
        n_run += 1
    print(f"read-format cases identical: {n_run}, skipped {skipped}", why[:12])
    assert n_run >= 60, (n_run, skipped, why)


def _prepare(tmp_path, case):
    """(R, Q, N, paired, local, sc, pe, ekw, index base) of a corpus case for the compiled engines, or None when its options are the caller's"""
    toks, kw, sc, pe_kw, local = _options(case)
    if any(k.startswith("_") for k in kw) or isinstance(sc.score_min_func, _BwaSwLike):
        return None                                        # trimming / skipping is the caller's; --bwa-sw-like's threshold is not in bt2g_policy_params
    paired = "mate1s" in case
    if paired and ("mate1fw" in case or "mate2fw" in case) and "pol" not in pe_kw:
        m1, m2 = case.get("mate1fw", 1), case.get("mate2fw", 0)
        if m1 == m2:
            pe_kw["pol"] = 1
        elif not m1:
            pe_kw["pol"] = 4
    base = _index_for(tmp_path, case["ref"])
    if not paired:
        seqs = case["reads"]
        quals = [(case.get("quals") or [None] * len(seqs))[k] or "I" * len(s) for k, s in enumerate(seqs)]
        names = [(case.get("names") or [None] * len(seqs))[k] or f"r{k}" for k in range(len(seqs))]
        R = [_codes(s) for s in seqs]
        Q = [np.frombuffer(q.encode(), dtype=np.uint8) for q in quals]
        N = names
    else:
        s1, s2 = case["mate1s"], case["mate2s"]
        q1 = [(case.get("qual1s") or [None] * len(s1))[k] or "I" * len(s) for k, s in enumerate(s1)]
        q2 = [(case.get("qual2s") or [None] * len(s2))[k] or "I" * len(s) for k, s in enumerate(s2)]
        names = [(case.get("names") or [None] * len(s1))[k] or f"r{k}" for k in range(len(s1))]
        R = [x for p in zip((_codes(s) for s in s1), (_codes(s) for s in s2)) for x in p]
        Q = [np.frombuffer(x.encode(), dtype=np.uint8) for p in zip(q1, q2) for x in p]
        N = [x for k in range(len(s1)) for x in (names[k] + "/1", names[k] + "/2")]
    if any(len(r) == 0 for r in R):
        return None
    ekw = dict(kw)
    ekw.setdefault("seed", 0)
    pe = policy.PairedEndPolicy(local=local, **pe_kw)
    return R, Q, N, paired, local, sc, pe, ekw, base


def _params(policy_params, paired, local, sc, pe, ekw):
    return policy_params("sensitive", local=local, paired=paired, seed=ekw.get("seed", 0), k=ekw.get("k"), all_hits=ekw.get("all_hits", False),
                         mhits=ekw.get("mhits", 50), nofw=ekw.get("nofw", False), norc=ekw.get("norc", False),
                         discord=ekw.get("discord", True), mixed=ekw.get("mixed", True), pe=pe, sc=sc, seed_len=ekw.get("seed_len"),
                         seed_rounds=ekw.get("seed_rounds"), dp_fail_streak=ekw.get("dp_fail_streak"), ival=ekw.get("ival"))


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_state_machine_on_the_regression_corpus(tmp_path):
    """EVERY usable corpus case within the device engine's reporting mode (-M) through csrc/xengine.cuh compiled for the host
    (bt2g_xengine_align_host) and through the coroutine engine (bt2g_policy_align), both over the C oracle table: identical result
    arrays -- odd scoring schemes, seed options, pairing options, tiny references and reads included"""
    from bowtie2_b200.lib import ReadBatch, load_library, policy_align, policy_params
    from oracle_lib import oracle_policy_table
    lib = load_library()
    n_sm = 0
    for ci, case in enumerate(_cases()):
        if not _usable(case):
            continue
        prep = _prepare(tmp_path, case)
        if prep is None:
            continue
        R, Q, N, paired, local, sc, pe, ekw, base = prep
        if ekw.get("k") is not None or ekw.get("all_hits", False) or max(len(r) for r in R) > 512:
            continue
        be, keep = oracle_policy_table(Oracle(base), local, 4, sc)
        prm = _params(policy_params, paired, local, sc, pe, ekw)
        res, ops, pairs, _ = policy_align(lib, be, prm, ReadBatch.from_list(R, Q), N)
        res2, ops2, pairs2, _ = policy_align(lib, be, prm, ReadBatch.from_list(R, Q), N, entry="bt2g_xengine_align_host")
        for f in ("found", "score", "score2", "fw", "tidx", "refoff", "nops", "trim_left", "trim_right", "mapq", "pad"):
            assert np.array_equal(res2[f], res[f]), (ci, case.get("name"), f, res2[f], res[f])
        assert np.array_equal(ops2, ops), (ci, case.get("name"))
        if paired:
            assert np.array_equal(pairs2["pair_type"], pairs["pair_type"]), (ci, case.get("name"))
        n_sm += 1
    assert n_sm >= 50, n_sm


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_compiled_engine_on_the_regression_corpus(tmp_path):
    """the same corpus through csrc/policy_engine.cpp (over the C oracle table): the primary alignment of every read / pair equals
    the Python engine's (which the test above pins against the reference program), options and odd scoring schemes included"""
    from bowtie2_b200.lib import ReadBatch, load_library, policy_align, policy_params
    from oracle_lib import oracle_policy_table
    from test_policy_engine_cpp import _python_results
    lib = load_library()
    n_run = n_sm = 0
    for ci, case in enumerate(_cases()):
        if ci % 2 or not _usable(case):                        # every other case: keeps the CPU suite short
            continue
        toks, kw, sc, pe_kw, local = _options(case)
        if any(k.startswith("_") for k in kw) or isinstance(sc.score_min_func, _BwaSwLike):
            continue                                           # trimming / skipping is the caller's; --bwa-sw-like's threshold is not in bt2g_policy_params
        paired = "mate1s" in case
        if paired and ("mate1fw" in case or "mate2fw" in case) and "pol" not in pe_kw:
            m1, m2 = case.get("mate1fw", 1), case.get("mate2fw", 0)
            if m1 == m2:
                pe_kw["pol"] = 1
            elif not m1:
                pe_kw["pol"] = 4
        base = _index_for(tmp_path, case["ref"])
        if not paired:
            seqs = case["reads"]
            quals = [(case.get("quals") or [None] * len(seqs))[k] or "I" * len(s) for k, s in enumerate(seqs)]
            names = [(case.get("names") or [None] * len(seqs))[k] or f"r{k}" for k in range(len(seqs))]
            R = [_codes(s) for s in seqs]
            Q = [np.frombuffer(q.encode(), dtype=np.uint8) for q in quals]
            N = names
        else:
            s1, s2 = case["mate1s"], case["mate2s"]
            q1 = [(case.get("qual1s") or [None] * len(s1))[k] or "I" * len(s) for k, s in enumerate(s1)]
            q2 = [(case.get("qual2s") or [None] * len(s2))[k] or "I" * len(s) for k, s in enumerate(s2)]
            names = [(case.get("names") or [None] * len(s1))[k] or f"r{k}" for k in range(len(s1))]
            R = [x for p in zip((_codes(s) for s in s1), (_codes(s) for s in s2)) for x in p]
            Q = [np.frombuffer(x.encode(), dtype=np.uint8) for p in zip(q1, q2) for x in p]
            N = [x for k in range(len(s1)) for x in (names[k] + "/1", names[k] + "/2")]
        if any(len(r) == 0 for r in R):
            continue
        ekw = dict(kw)
        ekw.setdefault("seed", 0)
        pe = policy.PairedEndPolicy(local=local, **pe_kw)
        want_res, want_ops, want_pairs = _python_results(base, R, Q, N, paired, local, 4, sc=sc, scoring=sc,
                                                         **({"pe": pe} if paired else {}),
                                                         **{k: v for k, v in ekw.items() if paired or k not in ("mixed", "discord")})
        be, keep = oracle_policy_table(Oracle(base), local, 4, sc)
        prm = policy_params("sensitive", local=local, paired=paired, seed=ekw.get("seed", 0), k=ekw.get("k"), all_hits=ekw.get("all_hits", False),
                            mhits=ekw.get("mhits", 50), nofw=ekw.get("nofw", False), norc=ekw.get("norc", False),
                            discord=ekw.get("discord", True), mixed=ekw.get("mixed", True), pe=pe, sc=sc, seed_len=ekw.get("seed_len"),
                            seed_rounds=ekw.get("seed_rounds"), dp_fail_streak=ekw.get("dp_fail_streak"), ival=ekw.get("ival"))
        res, ops, pairs, _ = policy_align(lib, be, prm, ReadBatch.from_list(R, Q), N)
        for f in ("found", "score", "score2", "fw", "tidx", "refoff", "nops", "trim_left", "trim_right", "mapq", "pad"):
            assert np.array_equal(res[f], want_res[f]), (ci, case.get("name"), f, res[f], want_res[f])
        assert np.array_equal(ops[:, :want_ops.shape[1]], want_ops), (ci, case.get("name"))
        if paired:
            assert np.array_equal(pairs["pair_type"], want_pairs["pair_type"]), (ci, case.get("name"))
        n_run += 1
        # ... and through the device engine's state machine compiled for the host (csrc/xengine.cuh via bt2g_xengine_align_host), for the
        # reporting mode it serves (-M; -k / -a go through the coroutine engine): the same records
        if ekw.get("k") is None and not ekw.get("all_hits", False) and max(len(r) for r in R) <= 512:
            res2, ops2, pairs2, _ = policy_align(lib, be, prm, ReadBatch.from_list(R, Q), N, entry="bt2g_xengine_align_host")
            for f in ("found", "score", "score2", "fw", "tidx", "refoff", "nops", "trim_left", "trim_right", "mapq", "pad"):
                assert np.array_equal(res2[f], res[f]), (ci, case.get("name"), "state machine", f, res2[f], res[f])
            assert np.array_equal(ops2, ops), (ci, case.get("name"), "state machine")
            if paired:
                assert np.array_equal(pairs2["pair_type"], pairs["pair_type"]), (ci, case.get("name"), "state machine")
            n_sm += 1
    assert n_run >= 55 and n_sm >= 25, (n_run, n_sm)
