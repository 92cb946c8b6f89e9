"""Python model of the optimised Poseidon schedule the CUDA kernels run.

Test helper (lives beside the tests, imports the oracle): derives, from the
reference's (C, M), the tables the library derives in C++ at init
(infimum_b200/csrc/host_params.cpp) and evaluates the optimised schedule on
Python ints.  tests/test_opt_model.py checks model == dense oracle, and
tests/test_capi_tables.py checks the library's tables == this model's.

Schedule (exact field identities, so still bit-identical to
pallet/src/hash/poseidon.rs:184-203):

  s      = [tag, in...] + C_0
  r=0..2 : s = M . sbox(s) + C_{r+1}
  r=3    : s = PRE . sbox(s) + k_0 e0                     PRE = B_0 . M
  j=0..RP-1 (partial): s0 = s0^5 ; s = S_j . s ; s0 += k_{j+1}   (last: s += D)
  r=4+RP..7+RP : s = M . sbox(s) + C_{r+1}                (C_{8+RP} = 0)
  out = s[0]

k_j / D: round constants of the partial section pushed forward through M
(valid because the partial S-box is the identity on elements 1..t-1).
S_j = [[m00, v.Mhat^-1],[w, I]] and B_j = diag(1, Mhat) come from factoring
N = S.B right to left through the partial rounds, each B commuting with the
preceding partial S-box and merging into the previous round's matrix.
"""
from oracle.poseidon_ref import P, poseidon_parameters


def _matmul(A, B):
    n, m, k = len(A), len(B[0]), len(B)
    return [[sum(A[i][x] * B[x][j] for x in range(k)) % P for j in range(m)] for i in range(n)]


def _matvec(A, v):
    return [sum(a * x for a, x in zip(row, v)) % P for row in A]


def _inv(A):
    n = len(A)
    M = [list(r) + [1 if i == j else 0 for j in range(n)] for i, r in enumerate(A)]
    for c in range(n):
        piv = next(r for r in range(c, n) if M[r][c] % P)
        M[c], M[piv] = M[piv], M[c]
        iv = pow(M[c][c], P - 2, P)
        M[c] = [x * iv % P for x in M[c]]
        for r in range(n):
            if r != c and M[r][c]:
                f = M[r][c]
                M[r] = [(x - f * y) % P for x, y in zip(M[r], M[c])]
    return [r[n:] for r in M]


def derive(t):
    ark, mds, rf, rp = poseidon_parameters(t)
    half = rf // 2
    C = [list(ark[r * t:(r + 1) * t]) for r in range(rf + rp)]
    M = [list(r) for r in mds]

    # --- constant compression through the partial section -------------------
    carry = [0] * t
    k = []
    for j in range(rp):
        c = [(a + b) % P for a, b in zip(C[half + j], carry)]
        k.append(c[0])
        c[0] = 0
        carry = _matvec(M, c)
    D = [(a + b) % P for a, b in zip(C[half + rp], carry)]

    # --- sparse factorisation, last partial round first ----------------------
    sparse = [None] * rp
    N = M
    for j in reversed(range(rp)):
        m00 = N[0][0]
        v = N[0][1:]
        w = [N[i][0] for i in range(1, t)]
        Nhat = [row[1:] for row in N[1:]]
        Ninv = _inv(Nhat)
        vp = [sum(v[x] * Ninv[x][c] for x in range(t - 1)) % P for c in range(t - 1)]
        sparse[j] = ([m00] + vp, w)                       # row 0, column 0 below the corner
        B = [[1] + [0] * (t - 1)] + [[0] + Nhat[i] for i in range(t - 1)]
        # B commutes with this round's partial S-box and "+k e0", so it slides
        # back to just after the previous round's matrix: that round now
        # applies M first and B second, i.e. the product B.M
        N = _matmul(B, M)
    PRE = N

    # --- unit leading coefficient: carry s0 as u = s0 / lambda_j --------------------
    # lambda_0 = 1, lambda_{j+1} = m00_j * lambda_j^5.  With z = u^5 (so s0^5 = lambda_j^5 z)
    #     u'   = z + (v_j / lambda_{j+1}) . s[1:] + kv_j / lambda_{j+1}
    #     s_i' = s_i + (w_j[i] * lambda_j^5) * z
    # and after the last partial round s0 = lambda_RP * u, which the next full round
    # only sees through s0^5 = lambda_RP^5 u^5: column 0 of that round's matrix is
    # scaled instead (TAIL0).  One product less per partial round, and the S-box
    # chain u -> u^2 -> u^4 -> u^5 -> u' has no multiplication after the S-box.
    kv = lambda j: k[j + 1] if j + 1 < rp else D[0]
    lam = [1]
    scaled = []
    for j in range(rp):
        (row0, w) = sparse[j]
        l5 = pow(lam[j], 5, P)
        nxt = row0[0] * l5 % P
        assert nxt != 0
        inv_n = pow(nxt, P - 2, P)
        scaled.append(([x * inv_n % P for x in row0[1:]], [x * l5 % P for x in w], kv(j) * inv_n % P))
        lam.append(nxt)
    l5 = pow(lam[rp], 5, P)
    TAIL0 = [[M[i][0] * l5 % P] + list(M[i][1:]) for i in range(t)]
    return dict(t=t, rf=rf, rp=rp, C=C, M=M, PRE=PRE, k=k, D=D, sparse=sparse, lam=lam, scaled=scaled, TAIL0=TAIL0)


def hash_opt(inputs, tag=0, tables=None):
    t = len(inputs) + 1
    T = tables or derive(t)
    rp, M, C = T["rp"], T["M"], T["C"]
    sb = lambda x: pow(x, 5, P)
    s = [(a + b) % P for a, b in zip([tag % P] + [x % P for x in inputs], C[0])]
    for r in range(3):
        s = [(a + b) % P for a, b in zip(_matvec(M, [sb(x) for x in s]), C[r + 1])]
    s = _matvec(T["PRE"], [sb(x) for x in s])
    s[0] = (s[0] + T["k"][0]) % P
    for j in range(rp):
        row0, w = T["sparse"][j]
        x0 = sb(s[0])
        new0 = (row0[0] * x0 + sum(a * b for a, b in zip(row0[1:], s[1:]))) % P
        s = [new0] + [(s[i] + w[i - 1] * x0) % P for i in range(1, t)]
        if j + 1 < rp:
            s[0] = (s[0] + T["k"][j + 1]) % P
        else:
            s = [(a + b) % P for a, b in zip(s, T["D"])]
    for r in range(4 + rp, 8 + rp):
        s = _matvec(M, [sb(x) for x in s])
        if r + 1 < 8 + rp:
            s = [(a + b) % P for a, b in zip(s, C[r + 1])]
    return s[0]


def paired(t):
    """Widths whose kernels pair partial rounds (poseidon.cuh paired_rounds): all
    of 2..8; an odd round count leaves one ordinary round at the end."""
    return t >= 2


def hash_opt_paired(inputs, tag=0, tables=None):
    """Same schedule with the partial rounds taken two at a time, the way the
    kernels for widths >= 4 run them: s[1..] is updated once per pair and round
    B's row sees round A's update through the scalar c_B = row0_B[1..] . w_A."""
    t = len(inputs) + 1
    T = tables or derive(t)
    rp, M, C = T["rp"], T["M"], T["C"]
    sb = lambda x: pow(x, 5, P)
    s = [(a + b) % P for a, b in zip([tag % P] + [x % P for x in inputs], C[0])]
    for r in range(3):
        s = [(a + b) % P for a, b in zip(_matvec(M, [sb(x) for x in s]), C[r + 1])]
    s = _matvec(T["PRE"], [sb(x) for x in s])
    s[0] = (s[0] + T["k"][0]) % P
    kv = lambda j: T["k"][j + 1] if j + 1 < rp else T["D"][0]
    for jp in range(rp // 2):
        (rowA, wA), (rowB, wB) = T["sparse"][2 * jp], T["sparse"][2 * jp + 1]
        xa = sb(s[0])
        n = (rowA[0] * xa + sum(a * b for a, b in zip(rowA[1:], s[1:])) + kv(2 * jp)) % P
        xb = sb(n)
        cB = sum(a * b for a, b in zip(rowB[1:], wA)) % P
        s0 = (rowB[0] * xb + sum(a * b for a, b in zip(rowB[1:], s[1:])) + cB * xa + kv(2 * jp + 1)) % P
        s = [s0] + [(s[i] + wA[i - 1] * xa + wB[i - 1] * xb) % P for i in range(1, t)]
    if rp % 2:                                      # the odd round out, in the ordinary form
        row0, w = T["sparse"][rp - 1]
        x0 = sb(s[0])
        new0 = (row0[0] * x0 + sum(a * b for a, b in zip(row0[1:], s[1:])) + kv(rp - 1)) % P
        s = [new0] + [(s[i] + w[i - 1] * x0) % P for i in range(1, t)]
    s = [s[0]] + [(a + b) % P for a, b in zip(s[1:], T["D"][1:])]
    for r in range(4 + rp, 8 + rp):
        s = _matvec(M, [sb(x) for x in s])
        if r + 1 < 8 + rp:
            s = [(a + b) % P for a, b in zip(s, C[r + 1])]
    return s[0]


def hash_opt_scaled(inputs, tag=0, tables=None, paired_rounds=True):
    """The schedule the kernels run since round 2: partial rounds with the leading
    coefficient scaled to one (see derive), taken in pairs when `paired_rounds`."""
    t = len(inputs) + 1
    T = tables or derive(t)
    rp, M, C = T["rp"], T["M"], T["C"]
    sb = lambda x: pow(x, 5, P)
    s = [(a + b) % P for a, b in zip([tag % P] + [x % P for x in inputs], C[0])]
    for r in range(3):
        s = [(a + b) % P for a, b in zip(_matvec(M, [sb(x) for x in s]), C[r + 1])]
    s = _matvec(T["PRE"], [sb(x) for x in s])
    u = (s[0] + T["k"][0]) % P                              # lambda_0 = 1
    rest = s[1:]
    j = 0
    if paired_rounds:
        for jp in range(rp // 2):
            (vA, wA, kA), (vB, wB, kB) = T["scaled"][2 * jp], T["scaled"][2 * jp + 1]
            za = sb(u)
            n = (za + sum(a * b for a, b in zip(vA, rest)) + kA) % P
            zb = sb(n)
            cB = sum(a * b for a, b in zip(vB, wA)) % P
            u = (zb + sum(a * b for a, b in zip(vB, rest)) + cB * za + kB) % P
            rest = [(rest[i] + wA[i] * za + wB[i] * zb) % P for i in range(t - 1)]
        j = 2 * (rp // 2)
    for j in range(j, rp):
        v, w, kk = T["scaled"][j]
        z = sb(u)
        u = (z + sum(a * b for a, b in zip(v, rest)) + kk) % P
        rest = [(rest[i] + w[i] * z) % P for i in range(t - 1)]
    rest = [(a + b) % P for a, b in zip(rest, T["D"][1:])]
    # first round of the second half: S-box on u itself, lambda_RP^5 sits in column 0 of TAIL0
    s = _matvec(T["TAIL0"], [sb(u)] + [sb(x) for x in rest])
    s = [(a + b) % P for a, b in zip(s, C[4 + rp + 1])]
    for r in range(4 + rp + 1, 8 + rp):
        s = _matvec(M, [sb(x) for x in s])
        if r + 1 < 8 + rp:
            s = [(a + b) % P for a, b in zip(s, C[r + 1])]
    return s[0]


# ---------------------------------------------------------------------------------------------
# Functional basis for width 3 (round 2, Layout<3>::FB in poseidon.cuh)
# ---------------------------------------------------------------------------------------------
def derive_fb(t=3, tables=None):
    """Partial rounds of width 3 with the two passive state elements carried as the two
    linear functionals the NEXT pair of rounds reads, instead of as themselves.

    In the unit-leading form (derive) a pair of rounds A, B is
        z_a = u^5 ; n = z_a + v_A.rest + k_A ; z_b = n^5 ; u' = z_b + v_B.rest + c z_a + k_B
        rest' = rest + w_A z_a + w_B z_b                         (c = v_B.w_A)
    i.e. 9 products and 4 reductions.  `rest` is two-dimensional, so it is determined by
        a = v_A.rest + k_A ,  b = v_B.rest + k_B
    (the 2x2 matrix [v_A; v_B] is invertible for these parameters), and the pair becomes
        z_a = u^5 ; n = z_a + a ; z_b = n^5 ; u' = z_b + b + c z_a
        a' = ga.(a, b, z_a, z_b) + ka ,  b' = gb.(a, b, z_a, z_b) + kb
    with a', b' the functionals of the next pair: 9 products, 3 reductions.  The first (a, b)
    come out of the merged round-3 matrix (rows 1, 2 of PRE replaced by v_0.PRE[1:], v_1.PRE[1:]);
    an odd round count ends with one single round that reads a = v.rest + k and b = rest[1] and
    returns the plain state elements (constants D folded in) to the second half."""
    assert t == 3
    T = tables or derive(t)
    rp, sc, D, PRE = T["rp"], T["scaled"], T["D"], T["PRE"]
    n_pairs = rp // 2
    inv = lambda x: pow(x, P - 2, P)

    def functionals(j):
        """The two functionals (vector, constant) read when the section resumes at round j."""
        if j + 1 < rp:
            return (sc[j][0], sc[j][2]), (sc[j + 1][0], sc[j + 1][2])
        if j < rp:                                   # odd count: the single round's row, and rest[1] itself
            return (sc[j][0], sc[j][2]), ([0, 1], 0)
        return ([1, 0], D[1]), ([0, 1], D[2])        # even count: the state itself, constants of the tail folded in

    (f1, c1), (f2, c2) = functionals(0)
    pre = [list(PRE[0]),
           [(f1[0] * PRE[1][x] + f1[1] * PRE[2][x]) % P for x in range(3)],
           [(f2[0] * PRE[1][x] + f2[1] * PRE[2][x]) % P for x in range(3)]]
    pre_v = [T["k"][0], c1, c2]
    pairs = []
    for p in range(n_pairs):
        A, B = 2 * p, 2 * p + 1
        (vA, wA, kA), (vB, wB, kB) = sc[A], sc[B]
        det = (vA[0] * vB[1] - vA[1] * vB[0]) % P
        assert det != 0
        di = inv(det)
        Vi = [[vB[1] * di % P, -vA[1] * di % P], [-vB[0] * di % P, vA[0] * di % P]]   # [v_A; v_B]^-1
        c = (vB[0] * wA[0] + vB[1] * wA[1]) % P
        rows = []
        for f, const in functionals(A + 2):
            g = [(f[0] * Vi[0][x] + f[1] * Vi[1][x]) % P for x in range(2)]         # f . V^-1
            ga = [g[0], g[1], (f[0] * wA[0] + f[1] * wA[1]) % P, (f[0] * wB[0] + f[1] * wB[1]) % P]
            rows.append((ga, (const - g[0] * kA - g[1] * kB) % P))
        pairs.append((c, rows[0], rows[1]))
    last = None
    if rp % 2:
        v, w, k = sc[rp - 1]
        assert v[0] != 0
        i0 = inv(v[0])
        # rest_0 = (a - k - v_1 b) / v_0 ; s_1' = rest_0 + w_0 z + D_1 ; s_2' = b + w_1 z + D_2
        last = dict(g1=[i0, -v[1] * i0 % P, w[0]], k1=(D[1] - k * i0) % P, w2=w[1], d2=D[2])
    return dict(pre=pre, pre_v=pre_v, pairs=pairs, last=last)


def hash_opt_fb(inputs, tag=0, tables=None, fb=None):
    """Width 3 through the functional-basis partial rounds (derive_fb)."""
    t = len(inputs) + 1
    T = tables or derive(t)
    F = fb or derive_fb(t, T)
    rp, M, C = T["rp"], T["M"], T["C"]
    sb = lambda x: pow(x, 5, P)
    s = [(a + b) % P for a, b in zip([tag % P] + [x % P for x in inputs], C[0])]
    for r in range(3):
        s = [(a + b) % P for a, b in zip(_matvec(M, [sb(x) for x in s]), C[r + 1])]
    u, a, b = [(x + y) % P for x, y in zip(_matvec(F["pre"], [sb(x) for x in s]), F["pre_v"])]
    for c, (ga, ka), (gb, kb) in F["pairs"]:
        za = sb(u)
        n = (za + a) % P
        zb = sb(n)
        u = (zb + b + c * za) % P
        q = (a, b, za, zb)
        a, b = (sum(x * y for x, y in zip(ga, q)) + ka) % P, (sum(x * y for x, y in zip(gb, q)) + kb) % P
    if F["last"]:
        L = F["last"]
        z = sb(u)
        u = (z + a) % P
        a, b = (L["g1"][0] * a + L["g1"][1] * b + L["g1"][2] * z + L["k1"]) % P, (b + L["w2"] * z + L["d2"]) % P
    s = _matvec(T["TAIL0"], [sb(u), sb(a), sb(b)])
    s = [(x + y) % P for x, y in zip(s, C[4 + rp + 1])]
    for r in range(4 + rp + 1, 8 + rp):
        s = _matvec(M, [sb(x) for x in s])
        if r + 1 < 8 + rp:
            s = [(x + y) % P for x, y in zip(s, C[r + 1])]
    return s[0]


# ---------------------------------------------------------------------------------------------
# Width 3, rows over values that exist anyway (Layout<3>::FB, the form the kernels run)
# ---------------------------------------------------------------------------------------------
def derive_fb2(t=3, tables=None):
    """derive_fb with b eliminated.  In a pair of rounds (A, B)
        z_a = u^5 ; n = z_a + a ; z_b = n^5 ; u' = z_b + b + c z_a
    b is only ever ADDED, so it need not exist as a reduced value: with b = u' - z_b - c z_a the
    next pair's two rows become rows over Q = (a, u', z_a, z_b), four values that exist anyway,
        a'            = ha . Q + ka
        b' + c' z_a'  = hb . Q + c' z_a' + kb        (one five-term row, once z_a' is known)
    i.e. 9 products and 2 reductions per pair next to the two S-boxes.  The odd round (RP = 57)
    goes first, in the plain form: the merged round-3 matrix yields u_0 and F_i = v_i.rest + k_i
    (i = 1, 2), u_1 = z_0 + al F1 + be F2 + k (v_0 = al v_1 + be v_2), and the first pair reads
    Q = (F1, u_1, z_0, F2) through selector rows.  After the last pair two rows over Q return the
    plain state elements (tail constants D folded in)."""
    assert t == 3
    T = tables or derive(t)
    rp, sc, D, PRE = T["rp"], T["scaled"], T["D"], T["PRE"]
    assert rp % 2 == 1
    inv = lambda x: pow(x, P - 2, P)
    dotp = lambda x, y: sum(a * b for a, b in zip(x, y)) % P

    def inv2(r0, r1):
        di = inv((r0[0] * r1[1] - r0[1] * r1[0]) % P)
        return [[r1[1] * di % P, -r0[1] * di % P], [-r1[0] * di % P, r0[0] * di % P]]

    (v0, w0, k0), (v1, _, k1), (v2, _, k2) = sc[0], sc[1], sc[2]
    pre = [list(PRE[0]),
           [(v1[0] * PRE[1][x] + v1[1] * PRE[2][x]) % P for x in range(3)],
           [(v2[0] * PRE[1][x] + v2[1] * PRE[2][x]) % P for x in range(3)]]
    pre_v = [T["k"][0], k1, k2]
    Vi = inv2(v1, v2)
    ab = [(v0[0] * Vi[0][x] + v0[1] * Vi[1][x]) % P for x in range(2)]          # v_0 = al v_1 + be v_2
    entry = (ab, (k0 - ab[0] * k1 - ab[1] * k2) % P)
    n_pairs = rp // 2
    recs = []
    # pair 0 reads Q = (F1, u_1, z_0, F2): a_0 = F1 + (v_1.w_0) z_0 ; b_0 = F2 + (v_2.w_0) z_0
    ha, ka = [1, 0, dotp(v1, w0), 0], 0
    hb, kb = [0, 0, dotp(v2, w0), 1], 0
    for p in range(n_pairs):
        A, B = 2 * p + 1, 2 * p + 2
        (vA, wA, kA), (vB, wB, kB) = sc[A], sc[B]
        c = dotp(vB, wA)
        recs.append((ha, ka, hb, c, kb))
        Vi = inv2(vA, vB)
        if p + 1 < n_pairs:
            targets = [(sc[A + 2][0], sc[A + 2][2]), (sc[B + 2][0], sc[B + 2][2])]
        else:
            targets = [([1, 0], D[1]), ([0, 1], D[2])]
        rows = []
        for f, const in targets:
            g = [(f[0] * Vi[0][x] + f[1] * Vi[1][x]) % P for x in range(2)]
            # g.(a, b) + (f.w_A) z_a + (f.w_B) z_b + const - g.(k_A, k_B), with b = u' - z_b - c z_a
            rows.append(([g[0], g[1], (dotp(f, wA) - c * g[1]) % P, (dotp(f, wB) - g[1]) % P],
                         (const - g[0] * kA - g[1] * kB) % P))
        (ha, ka), (hb, kb) = rows
    return dict(pre=pre, pre_v=pre_v, entry=entry, pairs=recs, exit=((ha, ka), (hb, kb)))


def hash_opt_fb2(inputs, tag=0, tables=None, fb=None):
    t = len(inputs) + 1
    T = tables or derive(t)
    F = fb or derive_fb2(t, T)
    rp, M, C = T["rp"], T["M"], T["C"]
    sb = lambda x: pow(x, 5, P)
    dotp = lambda x, y: sum(a * b for a, b in zip(x, y)) % P
    s = [(a + b) % P for a, b in zip([tag % P] + [x % P for x in inputs], C[0])]
    for r in range(3):
        s = [(a + b) % P for a, b in zip(_matvec(M, [sb(x) for x in s]), C[r + 1])]
    u, f1, f2 = [(x + y) % P for x, y in zip(_matvec(F["pre"], [sb(x) for x in s]), F["pre_v"])]
    z0 = sb(u)
    u = (z0 + dotp(F["entry"][0], (f1, f2)) + F["entry"][1]) % P
    Q = (f1, u, z0, f2)
    for ha, ka, hb, c, kb in F["pairs"]:
        za = sb(Q[1])
        a = (dotp(ha, Q) + ka) % P
        tt = (dotp(hb, Q) + c * za + kb) % P
        n = (za + a) % P
        zb = sb(n)
        Q = (a, (zb + tt) % P, za, zb)
    (h1, k1), (h2, k2) = F["exit"]
    s1, s2 = (dotp(h1, Q) + k1) % P, (dotp(h2, Q) + k2) % P
    s = _matvec(T["TAIL0"], [sb(Q[1]), sb(s1), sb(s2)])
    s = [(x + y) % P for x, y in zip(s, C[4 + rp + 1])]
    for r in range(4 + rp + 1, 8 + rp):
        s = _matvec(M, [sb(x) for x in s])
        if r + 1 < 8 + rp:
            s = [(x + y) % P for x, y in zip(s, C[r + 1])]
    return s[0]


# ---------------------------------------------------------------------------------------------
# History recurrence (Layout<T>::HR): the partial rounds with NO state but the S-box inputs and
# outputs of the last t-1 rounds
# ---------------------------------------------------------------------------------------------
def derive_hr(t, tables=None):
    """The passive state rest (n = t-1 elements) is only ever read through row functionals, and the
    n equations  u_{j+1-i} = z_{j-i} + v_{j-i}.rest^(j-i) + k_{j-i}  (i = 1..n) determine it from
    S-box inputs u and outputs z that exist anyway as reduced values.  So every partial round is
        z_j = u_j^5 ;  u_{j+1} = z_j + sum_{i=1..n} (al_i u_{j+1-i} + be_i z_{j-i}) + const
    2n products and ONE reduction beside the S-box, with nothing else to update.  The first n rounds
    read the functionals F_j = v_j.rest^(0) + k_j straight out of the merged round-3 matrix (row j+1
    of PRE replaced by v_j.PRE[1:]) plus the few z that exist by then; after the last round n rows
    over the same history return the plain state elements (tail constants D folded in).
    Records: steady[j - n] = (al[n], be[n], const) for rounds j = n..RP-1, in the order
    (u_j, u_{j-1}, ..; z_{j-1}, z_{j-2}, ..); exit rows likewise at j = RP."""
    T = tables or derive(t)
    rp, sc, D, PRE = T["rp"], T["scaled"], T["D"], T["PRE"]
    n = t - 1
    dotp = lambda x, y: sum(a * b for a, b in zip(x, y)) % P
    pre = [list(PRE[0])] + [[dotp(sc[j][0], [PRE[1 + x][c] for x in range(n)]) for c in range(t)] for j in range(n)]
    pre_v = [T["k"][0]] + [sc[j][2] for j in range(n)]
    boot = [[dotp(sc[j][0], sc[i][1]) for i in range(j)] for j in range(n)]          # boot[j][i] = v_j . w_i

    def row(j, f, kf):
        A = [sc[j - i][0] for i in range(1, n + 1)]
        Ai = _inv(A)
        g = [sum(f[x] * Ai[x][i] for x in range(n)) % P for i in range(n)]           # f . A^-1
        al = g
        be = [(-g[l - 1] + sum(g[i - 1] * dotp(sc[j - i][0], sc[j - l][1]) for i in range(l, n + 1))) % P
              for l in range(1, n + 1)]
        const = (kf - sum(g[i - 1] * sc[j - i][2] for i in range(1, n + 1))) % P
        return al, be, const

    steady = [row(j, sc[j][0], sc[j][2]) for j in range(n, rp)]
    unit = lambda i: [1 if x == i else 0 for x in range(n)]
    exit_rows = [row(rp, unit(i), D[1 + i]) for i in range(n)]
    return dict(n=n, pre=pre, pre_v=pre_v, boot=boot, steady=steady, exit=exit_rows)


def hash_opt_hr(inputs, tag=0, tables=None, hr=None):
    t = len(inputs) + 1
    T = tables or derive(t)
    H = hr or derive_hr(t, T)
    rp, M, C, n = T["rp"], T["M"], T["C"], t - 1
    sb = lambda x: pow(x, 5, P)
    dotp = lambda x, y: sum(a * b for a, b in zip(x, y)) % P
    s = [(a + b) % P for a, b in zip([tag % P] + [x % P for x in inputs], C[0])]
    for r in range(3):
        s = [(a + b) % P for a, b in zip(_matvec(M, [sb(x) for x in s]), C[r + 1])]
    s = [(x + y) % P for x, y in zip(_matvec(H["pre"], [sb(x) for x in s]), H["pre_v"])]
    us, zs = [s[0]], []                         # us[j] = u_j, zs[j] = z_j
    for j in range(rp):
        zs.append(sb(us[j]))
        if j < n:
            v = (s[1 + j] + dotp(H["boot"][j], zs[:j])) % P
        else:
            al, be, const = H["steady"][j - n]
            v = (dotp(al, [us[j + 1 - i] for i in range(1, n + 1)]) + dotp(be, [zs[j - i] for i in range(1, n + 1)]) + const) % P
        us.append((zs[j] + v) % P)
    rest = [(dotp(al, [us[rp + 1 - i] for i in range(1, n + 1)]) + dotp(be, [zs[rp - i] for i in range(1, n + 1)]) + const) % P
            for al, be, const in H["exit"]]
    s = _matvec(T["TAIL0"], [sb(us[rp])] + [sb(x) for x in rest])
    s = [(x + y) % P for x, y in zip(s, C[4 + rp + 1])]
    for r in range(4 + rp + 1, 8 + rp):
        s = _matvec(M, [sb(x) for x in s])
        if r + 1 < 8 + rp:
            s = [(x + y) % P for x, y in zip(s, C[r + 1])]
    return s[0]
