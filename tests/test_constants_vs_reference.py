"""In-container only: the Grain-LFSR constants the oracle regenerates must equal
the literals in the reference's pallet/src/hash/parameters.rs element for
element (all 12 widths).  Skipped where /root/reference is absent (GPU box)."""
import os
import re

import pytest

from oracle import poseidon_ref as O

PARAMS = "/root/reference/pallet/src/hash/parameters.rs"


@pytest.mark.skipif(not os.path.exists(PARAMS), reason="reference checkout not present")
def test_grain_constants_equal_reference_literals():
    src = open(PARAMS).read()
    parts = re.split(r"else if (\d+) == t", src)
    seen = set()
    for k in range(1, len(parts), 2):
        t = int(parts[k])
        nums = re.findall(r"BigInteger256::new\(\[\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),?\s*\]\)", parts[k + 1])
        vals = [int(a) | int(b) << 64 | int(c) << 128 | int(d) << 192 for a, b, c, d in nums]
        ark, mds, rf, rp = O.poseidon_parameters(t)
        assert len(vals) == t * (rf + rp) + t * t
        assert tuple(vals[: len(ark)]) == ark
        assert [tuple(vals[len(ark) + i * t: len(ark) + (i + 1) * t]) for i in range(t)] == list(mds)
        seen.add(t)
    assert seen == set(range(2, 14))
