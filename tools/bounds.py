#!/usr/bin/env python3
"""Range analysis of the device arithmetic (infimum_b200/csrc/fr.cuh, poseidon.cuh).

Values are tracked as multiples of p.  rho = p / 2^256.  A lazy Montgomery
dot of terms (alpha_j p) x (beta_j p) plus V < p gives a result below
(rho * sum alpha_j beta_j + 1 + rho) p, and csub2p leaves max(2 + eps, x - 2).
The script walks the schedule for each width and prints the largest value ever
held, which must stay below 2^256 / p = 5.29.
"""
P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
RHO = P / 2 ** 256
LIM = 2 ** 256 / P
EPS = 2 ** 224 / P


def dot(terms, v=1.0):
    return RHO * (sum(a * b for a, b in terms) + v) + 1.0


def csub2p(x):
    return max(2 + EPS, x - 2)


def sbox(x, track):
    assert x < 2 ** 255 / P, x                               # mont_sqr needs 2a to fit 8 limbs
    x2 = dot([(x, x)], 0); track(x2)
    assert x2 < 2 ** 255 / P
    x4 = dot([(x2, x2)], 0); track(x4)
    x5 = dot([(x4, x)], 0); track(x5)
    return x5


def csub4p(x):
    return max(4 + EPS, x - 4)


def hr_rounds(t):
    return 2 <= t <= 6                                       # poseidon.cuh: hr_rounds


def run(t, rp):
    worst = [0.0]
    rs = csub2p if t > 2 else (lambda x: x)                  # poseidon.cuh: RS = T > 2

    def track(x):
        worst[0] = max(worst[0], x)
        assert x < LIM, (t, x)
        return x

    # absorb_raw: any 256-bit integer, two range steps, + C_0 (< p), one more
    s = [csub2p(track(csub2p(csub4p(LIM)) + 1.0))] * t
    for r in range(4):                                       # first half
        x = [sbox(v, track) for v in s]
        s = [rs(track(dot([(xi, 1.0) for xi in x]))) for _ in range(t)]
    if hr_rounds(t):
        # history recurrence: rows over the last n S-box inputs (u) and outputs (z)
        n = t - 1
        hr_range = (lambda x: csub2p(csub4p(x))) if n >= 5 else csub2p      # poseidon.cuh: hr_range
        us, zs = [s[0]], []
        for j in range(n):                                   # bootstrap rounds: F_j + rows over the z so far
            zs.append(sbox(us[-1], track))
            v = s[1 + j] if j == 0 else csub2p(track(track(dot([(z, 1.0) for z in zs[:j]], 0)) + s[1 + j]))
            us.append(csub2p(track(zs[-1] + v)))
        for j in range(n, rp):
            z = sbox(us[-1], track)
            v = hr_range(track(dot([(u, 1.0) for u in us[-n:]] + [(x, 1.0) for x in zs[-n:]])))
            zs.append(z)
            us.append(csub2p(track(z + v)))
        rest = [hr_range(track(dot([(u, 1.0) for u in us[-n:]] + [(x, 1.0) for x in zs[-n:]]))) for _ in range(n)]
        s = [us[-1]] + rest
        rp = 0                                               # nothing left for the paired form below
    ri = csub2p if t >= 5 else (lambda x: x)                 # poseidon.cuh: RI = T >= 5
    for j in range(rp // 2):                                 # paired partial rounds, unit leading coefficient
        za = sbox(s[0], track)
        n = csub2p(track(ri(track(dot([(si, 1.0) for si in s[1:]]))) + za))
        zb = sbox(n, track)
        n0 = csub2p(track(ri(track(dot([(si, 1.0) for si in s[1:]] + [(za, 1.0)]))) + zb))
        s = [n0] + [csub2p(track(si + track(dot([(za, 1.0), (zb, 1.0)], 0)))) for si in s[1:]]
    for j in range(rp % 2):                                  # the odd round out
        z = sbox(s[0], track)
        n0 = csub2p(track(ri(track(dot([(si, 1.0) for si in s[1:]]))) + z))
        s = [n0] + [csub2p(track(si + track(dot([(z, 1.0)], 0)))) for si in s[1:]]
    if not hr_rounds(t):
        s = [s[0]] + [csub2p(track(si + 1.0)) for si in s[1:]]
    for r in range(3):
        x = [sbox(v, track) for v in s]
        s = [rs(track(dot([(xi, 1.0) for xi in x]))) for _ in range(t)]
    x = [sbox(v, track) for v in s]
    out = track(dot([(xi, 1.0) for xi in x], 0))
    out = csub2p(out)
    assert out < 3.0                                          # two exact subtractions of p suffice
    return worst[0], out


if __name__ == "__main__":
    RP = {2: 56, 3: 57, 4: 56, 5: 60, 6: 60, 7: 63, 8: 64}
    print("limit 2^256/p = %.4f, rho = p/2^256 = %.4f" % (LIM, RHO))
    for t in range(2, 9):
        w, o = run(t, RP[t])
        print("t=%d: largest intermediate %.3f p, pre-canonical output < %.3f p" % (t, w, o))
