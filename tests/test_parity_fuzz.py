"""A few fixed seeds of tests/parity_fuzz.py in the CPU suite: random genome / reads / preset / option sets through the device engine's
state machine on the host, every SAM record identical to the unmodified reference program's (oracle/_ref)."""
import os

import pytest

import parity_fuzz


@pytest.mark.skipif(not os.path.exists(parity_fuzz.REF), reason="oracle/_ref is not built")
@pytest.mark.parametrize("seed", [11, 100])
def test_random_configurations_identical_to_reference(seed, tmp_path):
    seen = set()
    for k in range(16 if seed == 11 else 8):
        c = parity_fuzz.draw_case(seed, k)
        n, bad, first, st, desc = parity_fuzz.run_case(c, str(tmp_path), n_unpaired=150, n_pairs=100)
        assert bad == 0, (k, desc, first)
        seen.update(c["flags"])
    if seed == 100:
        assert {"--no-discordant", "--mp", "--score-min", "--n-ceil"}.issubset(seen)
    # (seed 100, case 2: --no-discordant, with scoring options, on pairs whose mates each aligned once without a concordant pair -- records that read YT:Z:DP
    # unless the formatter is told about the option, BT2G_SAM_NO_DISCORDANT)
