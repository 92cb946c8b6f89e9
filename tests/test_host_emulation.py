"""The device arithmetic (fr.cuh / poseidon.cuh), compiled for the host with
every PTX block replaced by its C body, against Python integers and the oracle.
This is how the kernel logic is unit-tested where there is no GPU."""
import ctypes
import random

import pytest

import __graft_entry__ as entry
from oracle import poseidon_ref as O
from tests import opt_model
from tests.util import EDGE_VALUES

P = O.P
R = 1 << 256
RINV = pow(R, -1, P)
U8 = ctypes.c_uint32 * 8


def limbs(x):
    return U8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def val(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


@pytest.fixture(scope="module")
def emu():
    entry.build()
    from infimum_b200 import build as b
    lib = ctypes.CDLL(b.HOSTEMU)
    lib.hostemu_overflow_count.restype = ctypes.c_ulonglong
    return lib


def test_mont_mul_ranges(emu):
    rng = random.Random(11)
    lim = 2 * P + (1 << 224)
    cases = [(a % lim, b % lim) for a in EDGE_VALUES for b in EDGE_VALUES]
    cases += [(rng.randrange(lim), rng.randrange(lim)) for _ in range(3000)]
    cases += [(rng.randrange(R), rng.randrange(P)) for _ in range(500)]      # raw 256-bit input x constant
    base = emu.hostemu_overflow_count()
    for a, b in cases:
        r = U8()
        emu.hostemu_mont_mul(limbs(a), limbs(b), r)
        assert val(r) % P == a * b * RINV % P
        assert val(r) < 2 * P + 1
    assert emu.hostemu_overflow_count() == base


@pytest.mark.parametrize("n", range(1, 9))
def test_lazy_dot(emu, n):
    rng = random.Random(100 + n)
    lim = 2 * P + (1 << 224)
    base = emu.hostemu_overflow_count()
    for it in range(200):
        A = [rng.randrange(lim) for _ in range(n)]
        B = [rng.randrange(P) for _ in range(n)]
        V = rng.randrange(P)
        if it < 10:                       # worst case for the range analysis
            A, B, V = [lim - 1] * n, [P - 1] * n, P - 1
        elif it < 20:                     # all-ones limbs stress the carry chains
            A = [(0xFFFFFFFF << (32 * (it % 7))) | 0xFFFFFFFF] * n
        aa = (ctypes.c_uint32 * (8 * n))(*[(x >> (32 * i)) & 0xFFFFFFFF for x in A for i in range(8)])
        bb = (ctypes.c_uint32 * (8 * n))(*[(x >> (32 * i)) & 0xFFFFFFFF for x in B for i in range(8)])
        r = U8()
        assert emu.hostemu_dot(n, aa, bb, limbs(V) if it % 2 == 0 else None, r) == 0
        exp = (sum(x * y for x, y in zip(A, B)) + (V if it % 2 == 0 else 0)) * RINV % P
        assert val(r) % P == exp
        assert val(r) < lim
    assert emu.hostemu_overflow_count() == base


def test_redc_and_range_steps(emu):
    rng = random.Random(5)
    for x in EDGE_VALUES + [rng.randrange(R) for _ in range(300)]:
        r = U8()
        emu.hostemu_redc(limbs(x), r)
        assert val(r) % P == x * RINV % P and val(r) <= P
        y = limbs(x)
        emu.hostemu_csub2p(y)
        assert val(y) % P == x % P
        assert val(y) < max(2 * P + (1 << 224), x - 2 * P + 1)
        if x < 2 * P:
            z = limbs(x)
            emu.hostemu_csub_p_exact(z)
            assert val(z) == (x - P if x >= P else x)


def test_range_step_of_4p_and_the_widest_recurrence_row(emu):
    """csub4p (absorb_raw, hr_range for five passive elements) and the ten-term row of the width-6 recurrence at
    the top of its range: five S-box inputs just under 2p + 2^224, five S-box outputs at 1.7 p, every constant
    p - 1, the additive constant p - 1 -- the tightest accumulator of the whole schedule (tools/bounds.py: 4.59 p
    of 5.29 p).  No overflow, exact value mod p, below 2p + 2^224 after its range steps."""
    rng = random.Random(6)
    for x in EDGE_VALUES + [rng.randrange(R) for _ in range(300)] + [4 * P - 1, 4 * P, 4 * P + (1 << 224), 4 * P + (1 << 225)]:
        y = limbs(x)
        emu.hostemu_csub4p(y)
        assert val(y) % P == x % P
        assert val(y) < max(4 * P + (1 << 224), x - 4 * P + 1)
        emu.hostemu_csub2p(y)
        assert val(y) < 2 * P + (1 << 224)                  # any 256-bit integer: what absorb_raw relies on
    base = emu.hostemu_overflow_count()
    U80 = ctypes.c_uint32 * 80
    pack = lambda xs: U80(*[(x >> (32 * i)) & 0xFFFFFFFF for x in xs for i in range(8)])
    hi_u, hi_z = 2 * P + (1 << 224) - 1, 17 * P // 10
    cases = [([hi_u] * 5 + [hi_z] * 5, [P - 1] * 10, P - 1)]
    for _ in range(40):
        cases.append(([rng.randrange(2 * P + (1 << 224)) for _ in range(5)] + [rng.randrange(hi_z) for _ in range(5)],
                      [rng.randrange(P) for _ in range(10)], rng.randrange(P)))
    for a, b, v in cases:
        r = U8()
        # V enters the accumulator as is and is divided by R with everything else
        assert emu.hostemu_dot(10, pack(a), pack(b), limbs(v), r) == 0
        exp = (sum(x * y for x, y in zip(a, b)) + v) * RINV % P
        got = val(r)
        assert got % P == exp
        assert got < max(2 * P + (1 << 224), int(4.6 * P) - 2 * P)      # dot applies one step of 2p itself
        emu.hostemu_hr_range(5, r)
        assert val(r) % P == exp and val(r) < 2 * P + (1 << 224)
    assert emu.hostemu_overflow_count() == base


@pytest.mark.parametrize("t", range(2, 9))
def test_optimised_hash_equals_oracle(emu, t):
    rng = random.Random(200 + t)
    base = emu.hostemu_overflow_count()
    cases = [[1] * (t - 1), [0] * (t - 1), [R - 1] * (t - 1), [P] * (t - 1), [P - 1] * (t - 1)]
    cases += [[rng.choice(EDGE_VALUES) for _ in range(t - 1)] for _ in range(4)]
    cases += [[rng.randrange(R) for _ in range(t - 1)] for _ in range(6)]
    for i, ins in enumerate(cases):
        tag = None if i % 3 else rng.randrange(R)
        exp = O.poseidon_permute_hash([x % P for x in ins], (tag or 0) % P)
        for le, order in ((0, "big"), (1, "little")):
            buf = b"".join(x.to_bytes(32, order) for x in ins)
            out = ctypes.create_string_buffer(32)
            tagb = tag.to_bytes(32, order) if tag is not None else None
            assert emu.hostemu_hash(t, buf, tagb, out, le) == 0
            assert int.from_bytes(out.raw, order) == exp
    assert emu.hostemu_overflow_count() == base


@pytest.mark.parametrize("t", range(2, 14))
def test_library_grain_constants_equal_oracle(emu, t):
    ark, mds, rf, rp = O.poseidon_parameters(t)
    n = len(ark) + t * t
    buf = (ctypes.c_uint32 * (8 * n))()
    assert emu.hostemu_dense_params(t, buf) == n
    got = [sum(int(buf[8 * k + i]) << (32 * i) for i in range(8)) for k in range(n)]
    assert got[: len(ark)] == list(ark)
    assert got[len(ark):] == [x for row in mds for x in row]


@pytest.mark.parametrize("t", range(2, 9))
def test_library_optimised_tables_equal_python_derivation(emu, t):
    """Every entry of the table the library derives at init (host_params.cpp, Layout<T> in poseidon.cuh)
    against the independent Python derivation, in table order."""
    T = opt_model.derive(t)
    words = emu.hostemu_opt_table_words(t)
    buf = (ctypes.c_uint32 * words)()
    assert emu.hostemu_opt_table(t, buf) == words
    el = lambda k: sum(int(buf[8 * k + i]) << (32 * i) for i in range(8))
    rp, n = T["rp"], t - 1
    mont = lambda x: x * R % P
    vform = lambda x: x * R * R % P
    pos = [0]

    def expect(values, form=mont):
        for x in values:
            assert el(pos[0]) == form(x), pos[0]
            pos[0] += 1

    flat = lambda rows: [x for row in rows for x in row]
    ident = lambda x: x
    # ---- every schedule
    expect([1], lambda _: R * R % P)
    expect(T["C"][0], vform)
    expect([T["C"][0][0]])
    expect(flat(T["M"]))
    expect(flat(T["TAIL0"]))
    for r in range(3):
        expect(T["C"][r + 1], vform)
    for r in range(3):
        expect(T["C"][4 + rp + r + 1], vform)
    expect(T["M"][0], ident)
    expect(T["M"][0])
    # round 0 on unconverted inputs: X0 = (C_0[0])^5 / R^4, R0_M = M R^6, IN_C = C_0
    expect([pow(T["C"][0][0], 5, P) * pow(R, -4, P) % P], ident)
    expect(flat(T["M"]), lambda x: x * pow(R, 6, P) % P)
    expect(T["C"][0], ident)
    # ---- history recurrence (widths 2..6)
    if t <= 6:
        H = opt_model.derive_hr(t, T)
        expect(flat(H["pre"]))
        expect(H["pre_v"], vform)
        for j in range(1, n):
            expect(reversed(H["boot"][j]))                   # newest z first
        for al, be, const in H["steady"] + H["exit"]:
            expect(al)
            expect(be)
            expect([const], vform)
    # ---- paired sparse rounds
    expect(flat(T["PRE"]))
    expect([T["k"][0]], vform)
    expect([0] * (t - 1), ident)
    n_pairs = rp // 2 if opt_model.paired(t) else 0
    for jp in range(n_pairs):
        (vA, wA, kA), (vB, wB, kB) = T["scaled"][2 * jp], T["scaled"][2 * jp + 1]
        expect(vA)
        expect([kA], vform)
        expect(vB)
        expect([sum(a * b for a, b in zip(vB, wA)) % P])
        expect([kB], vform)
        expect([x for pair in zip(wA, wB) for x in pair])
    for j in range(2 * n_pairs, rp):
        v, w, kk = T["scaled"][j]
        expect(v)
        expect(w)
        expect([kk], vform)
    expect(T["D"][1:])
    expect([0 if j == 0 else sum(a * b for a, b in zip(T["scaled"][j][0], T["scaled"][j - 1][1])) % P for j in range(rp)])
    assert pos[0] * 8 == words


def test_mont_sqr_equals_mul(emu):
    rng = random.Random(77)
    lim = 2 * P + (1 << 224)
    cases = [x % lim for x in EDGE_VALUES] + [rng.randrange(lim) for _ in range(3000)]
    cases += [(1 << 255) - 1, 0x7FFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF,
              int("80000000" * 7 + "00000001", 16) >> 1, int("ffffffff" * 7, 16), int("80000000" * 8, 16) >> 1]
    base = emu.hostemu_overflow_count()
    for a in cases:
        assert a < (1 << 255)
        r, q = U8(), U8()
        emu.hostemu_mont_sqr(limbs(a), r)
        emu.hostemu_mont_mul(limbs(a), limbs(a), q)
        assert val(r) == val(q), hex(a)
        assert val(r) % P == a * a * RINV % P
    assert emu.hostemu_overflow_count() == base


@pytest.mark.parametrize("t", range(2, 9))
def test_warp_cooperative_schedule_equals_oracle(emu, t):
    """coop.cuh — the schedule of the kernel that takes the tree levels near the root —
    run with one host thread per role over barriers that follow the PTX named-barrier
    rules: same hash as the oracle, no barrier entered twice in a generation, no value
    past 2^256, whichever role is the slow one."""
    rng = random.Random(900 + t)
    base = emu.hostemu_overflow_count()
    cases = [[1] * (t - 1), [R - 1] * (t - 1), [P - 1] * (t - 1)] + [[rng.randrange(R) for _ in range(t - 1)] for _ in range(3)]
    jitters = [0, 1, 1 << t, (1 << t) - 2, 2, (1 << (t + 1)) - 1 - 1]
    for i, ins in enumerate(cases):
        exp = O.poseidon_permute_hash([x % P for x in ins], 0)
        buf = b"".join(x.to_bytes(32, "big") for x in ins)
        out = ctypes.create_string_buffer(32)
        assert emu.hostemu_coop_hash(t, buf, out, jitters[i % len(jitters)]) == 0
        assert int.from_bytes(out.raw, "big") == exp, (t, i)
    assert emu.hostemu_overflow_count() == base
