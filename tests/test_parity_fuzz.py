"""A few fixed seeds of tests/parity_fuzz.py in the CPU suite: random genome / reads / preset / option sets through the device engine's
state machine on the host, every SAM record identical to the unmodified reference program's (oracle/_ref)."""
import os

import numpy as np
import pytest

import parity_fuzz


@pytest.mark.skipif(not os.path.exists(parity_fuzz.REF), reason="oracle/_ref is not built")
@pytest.mark.parametrize("seed", [11, 105])
def test_random_configurations_identical_to_reference(seed, tmp_path):
    rng = np.random.default_rng(seed)
    seen = set()
    for k in range(16 if seed == 11 else 8):
        c = parity_fuzz.draw_case(rng)
        n, bad, first, st, desc = parity_fuzz.run_case(c, str(tmp_path), n_unpaired=150, n_pairs=100)
        assert bad == 0, (k, desc, first)
        seen.update(c["flags"])
    if seed == 105:
        assert {"--no-discordant", "--mp", "--rdg", "--n-ceil"}.issubset(seen)
    # (seed 105, case 2: --no-discordant, with scoring options, on pairs whose mates each aligned once without a concordant pair -- records that read YT:Z:DP
    # unless the formatter is told about the option, BT2G_SAM_NO_DISCORDANT)
