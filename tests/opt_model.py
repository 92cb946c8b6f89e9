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
