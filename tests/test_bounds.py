"""Range analysis of the shipped schedules (tools/bounds.py): no intermediate of any width reaches 2^256,
every S-box input stays below 2^255 (the squaring's precondition), the pre-canonical output needs at
most two exact subtractions.  The host-emulation tests count overflows on sampled inputs; this walks
the worst case."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import bounds  # noqa: E402


def test_every_width_stays_in_range():
    rp = {2: 56, 3: 57, 4: 56, 5: 60, 6: 60, 7: 63, 8: 64}
    for t in range(2, 9):
        worst, out = bounds.run(t, rp[t])          # asserts inside: < 2^256 everywhere, < 2^255 into every squaring
        assert worst < bounds.LIM and out < 3.0, (t, worst, out)
    # the widest row of the history recurrence (t = 6: ten terms) is the tightest spot
    assert 4.5 < bounds.run(6, 60)[0] < 4.7


def test_recurrence_rows_of_widths_7_and_8_are_out_of_range():
    """Why hr_rounds() stops at width 6: a 2n-term row of u < 2p and z < 1.7p is below (0.189 * 3.7 n + 1) p:
    4.50 p for n = 5, 5.20 p for n = 6 (within 2 % of 2^256 = 5.29 p: no margin for the eps terms), 5.90 p for n = 7."""
    row = lambda n: bounds.RHO * (2.0 + 1.7) * n + 1.0
    assert row(5) < bounds.LIM - 0.7
    assert bounds.LIM - 0.1 < row(6) < bounds.LIM
    assert row(7) > bounds.LIM
