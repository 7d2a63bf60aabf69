"""A few fixed cases of tests/parity_fuzz.py in the CPU suite: random genome / reads / preset / option sets through the device engine's
state machine on the host (and, for -k / -a, through the coroutine engine), every SAM record identical to the unmodified reference
program's (oracle/_ref)."""
import os

import pytest

import parity_fuzz

CASES = [
    (11, list(range(16))),
    # 100/2: --no-discordant on pairs whose mates each aligned once without a concordant pair (records that read YT:Z:DP unless the
    # formatter is told about the option, BT2G_SAM_NO_DISCORDANT)
    (100, list(range(8))),
    # 101/11, 44, 55: paired -k with an unaligned mate 1 beside a multiply aligned mate 2 (the unaligned record is the pair's LAST one);
    # 101/13: cheap gaps and a high match bonus -- an alignment with more edit ops than read length + 64
    (101, [11, 13, 44, 55]),
    # 92/180: 2-5 bp reads with one N under --n-ceil
    (92, [180]),
    # (not in the suite, 100 s of oracle DP: `python tests/parity_fuzz.py`'s case 64/182, an alignment with more ops than the coroutine
    # engine's first retry buffers held -- parity_fuzz.run_case(parity_fuzz.draw_case(64, 182), workdir))
]


@pytest.mark.skipif(not os.path.exists(parity_fuzz.REF), reason="oracle/_ref is not built")
@pytest.mark.parametrize("seed,cases", CASES + [(131, [3, 60, 74, 75]), (151, [10, 34])])
def test_random_configurations_identical_to_reference(seed, cases, tmp_path, monkeypatch):
    if seed in (131, 151):
        # (151/10, 34: pairs with an EMPTY mate 2 -- an unpaired read for the reference, the stream's solo engine)
        # as FILES too: stream.align_files_stream (reader, parser, engine, formatter, ordered writer, alignment summary) on the same cases --
        # ragged reads down to empty ones, empty mate 1, .bt2l, pairing options
        monkeypatch.setenv("BT2G_FUZZ_FILES", "1")
    seen = set()
    for k in cases:
        c = parity_fuzz.draw_case(seed, k)
        small = len(cases) > 4
        n, bad, first, st, desc = parity_fuzz.run_case(c, str(tmp_path), n_unpaired=150 if small else 300, n_pairs=100 if small else 200)
        assert bad == 0, (k, desc, first)
        seen.update(c["flags"])
    if seed == 100:
        assert {"--no-discordant", "--mp", "--score-min", "--n-ceil"}.issubset(seen)
    if seed == 101:
        assert "-k" in seen
