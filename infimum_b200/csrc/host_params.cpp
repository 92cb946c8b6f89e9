// See host_params.h.
#include "host_params.h"

#include <map>
#include <mutex>
#include <stdexcept>

#include "poseidon.cuh"

namespace inf {
namespace host {

// ---------------------------------------------------------------------------
// Grain LFSR (80 bit, self-shrinking output) — Poseidon paper, appendix on
// round-constant generation.
// ---------------------------------------------------------------------------
namespace {
struct Grain {
    uint8_t s[80];
    int head = 0;   // index of the oldest bit
    Grain(int field_bits, int t, int rf, int rp) {
        int n = 0;
        auto put = [&](unsigned v, int w) {
            for (int k = w - 1; k >= 0; k--) s[n++] = (v >> k) & 1;
        };
        put(1, 2);            // prime field
        put(0, 4);            // S-box x^alpha
        put(field_bits, 12);
        put(t, 12);
        put(rf, 10);
        put(rp, 10);
        for (int k = 0; k < 30; k++) s[n++] = 1;
        for (int k = 0; k < 160; k++) clock();
    }
    inline uint8_t at(int i) const { return s[(head + i) % 80]; }
    uint8_t clock() {
        uint8_t b = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
        s[head] = b;                 // overwrite oldest = append newest
        head = (head + 1) % 80;
        return b;
    }
    uint8_t bit() {
        for (;;) {
            uint8_t keep = clock();
            uint8_t out = clock();
            if (keep) return out;
        }
    }
    F next254() {   // 254 bits, most significant first
        F r = zero();
        for (int k = 253; k >= 0; k--)
            if (bit()) r.l[k / 64] |= 1ull << (k % 64);
        return r;
    }
};
}  // namespace

const DenseParams& grain_params(int t) {
    static std::map<int, DenseParams> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(t);
    if (it != cache.end()) return it->second;
    if (t < 2 || t > 13) throw std::invalid_argument("poseidon width out of range");
    DenseParams d;
    d.t = t;
    d.rf = 8;
    d.rp = partial_rounds(t);
    Grain g254(254, t, d.rf, d.rp);
    d.ark.reserve((d.rf + d.rp) * t);
    for (int k = 0; k < (d.rf + d.rp) * t; k++) {
        F v = g254.next254();
        while (geq(v, MOD)) v = g254.next254();      // rejection sampling
        d.ark.push_back(v);
    }
    // Cauchy MDS from 2t further field elements (reduced, not rejected)
    std::vector<F> xy;
    for (;;) {
        xy.clear();
        for (int k = 0; k < 2 * t; k++) xy.push_back(reduce256(g254.next254()));
        bool distinct = true;
        for (int a = 0; a < 2 * t && distinct; a++)
            for (int b = a + 1; b < 2 * t; b++)
                if (xy[a] == xy[b]) { distinct = false; break; }
        if (distinct) break;
    }
    d.mds.resize(t * t);
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) d.mds[i * t + j] = inv(add(xy[i], xy[t + j]));
    return cache.emplace(t, std::move(d)).first->second;
}

// ---------------------------------------------------------------------------
// small dense linear algebra over Fr
// ---------------------------------------------------------------------------
namespace {
typedef std::vector<F> Mat;   // n*n row-major

Mat matmul(const Mat& a, const Mat& b, int n) {
    Mat r(n * n, zero());
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            F acc = zero();
            for (int k = 0; k < n; k++) acc = add(acc, mul(a[i * n + k], b[k * n + j]));
            r[i * n + j] = acc;
        }
    return r;
}
std::vector<F> matvec(const Mat& a, const std::vector<F>& v, int n) {
    std::vector<F> r(n, zero());
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n; k++) r[i] = add(r[i], mul(a[i * n + k], v[k]));
    return r;
}
Mat matinv(Mat a, int n) {
    Mat r(n * n, zero());
    for (int i = 0; i < n; i++) r[i * n + i] = one();
    for (int c = 0; c < n; c++) {
        int piv = c;
        while (piv < n && a[piv * n + c].is_zero()) piv++;
        if (piv == n) throw std::runtime_error("singular matrix in sparse factorisation");
        if (piv != c)
            for (int k = 0; k < n; k++) {
                std::swap(a[piv * n + k], a[c * n + k]);
                std::swap(r[piv * n + k], r[c * n + k]);
            }
        F iv = inv(a[c * n + c]);
        for (int k = 0; k < n; k++) {
            a[c * n + k] = mul(a[c * n + k], iv);
            r[c * n + k] = mul(r[c * n + k], iv);
        }
        for (int i = 0; i < n; i++) {
            if (i == c || a[i * n + c].is_zero()) continue;
            F f = a[i * n + c];
            for (int k = 0; k < n; k++) {
                a[i * n + k] = sub(a[i * n + k], mul(f, a[c * n + k]));
                r[i * n + k] = sub(r[i * n + k], mul(f, r[c * n + k]));
            }
        }
    }
    return r;
}

inline void put_canon(std::vector<uint32_t>& tbl, int idx, const F& x) { to_limbs32(x, &tbl[idx * 8]); }
inline void put_mont(std::vector<uint32_t>& tbl, int idx, const F& x) { to_limbs32(to_mont(x), &tbl[idx * 8]); }
// x * R^2 mod p: value that becomes x*R after the accumulator's division by R
inline void put_v(std::vector<uint32_t>& tbl, int idx, const F& x) {
    to_limbs32(to_mont(to_mont(x)), &tbl[idx * 8]);
}

template <int T>
std::vector<uint32_t> build_opt(void) {
    using L = Layout<T>;
    const DenseParams& d = grain_params(T);
    const int rp = d.rp, half = d.rf / 2;
    std::vector<uint32_t> tbl(L::WORDS, 0u);
    auto C = [&](int r, int i) -> const F& { return d.ark[r * T + i]; };
    const Mat& M = d.mds;

    // constants of the partial section pushed forward (see tests/opt_model.py)
    std::vector<F> carry(T, zero()), k(rp), D(T);
    for (int j = 0; j < rp; j++) {
        std::vector<F> c(T);
        for (int i = 0; i < T; i++) c[i] = add(C(half + j, i), carry[i]);
        k[j] = c[0];
        c[0] = zero();
        carry = matvec(M, c, T);
    }
    for (int i = 0; i < T; i++) D[i] = add(C(half + rp, i), carry[i]);

    // sparse factorisation, last partial round first
    std::vector<std::vector<F>> row0(rp), wcol(rp);
    Mat N = M;
    for (int j = rp - 1; j >= 0; j--) {
        const int n = T - 1;
        Mat Nhat(n * n);
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) Nhat[a * n + b] = N[(a + 1) * T + (b + 1)];
        Mat Ninv = matinv(Nhat, n);
        row0[j].resize(T);
        row0[j][0] = N[0];
        for (int c = 0; c < n; c++) {
            F acc = zero();
            for (int x = 0; x < n; x++) acc = add(acc, mul(N[0 * T + (x + 1)], Ninv[x * n + c]));
            row0[j][c + 1] = acc;
        }
        wcol[j].resize(n);
        for (int a = 0; a < n; a++) wcol[j][a] = N[(a + 1) * T + 0];
        Mat B(T * T, zero());
        B[0] = one();
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) B[(a + 1) * T + (b + 1)] = Nhat[a * n + b];
        N = matmul(B, M, T);
    }
    const Mat& PRE = N;

    // unit leading coefficient (Layout<T> in poseidon.cuh, derive() in tests/opt_model.py):
    // lambda_0 = 1, lambda_{j+1} = m00_j lambda_j^5; v' = v / lambda_{j+1}, k' = k / lambda_{j+1},
    // w' = w lambda_j^5
    auto kv = [&](int j) -> const F& { return j + 1 < rp ? k[j + 1] : D[0]; };   // constant folded into round j's row
    auto pow5 = [](const F& x) { F x2 = mul(x, x); return mul(mul(x2, x2), x); };
    std::vector<std::vector<F>> vs(rp), ws(rp);
    std::vector<F> ks(rp);
    F lam = one();
    for (int j = 0; j < rp; j++) {
        const F l5 = pow5(lam);
        const F nxt = mul(row0[j][0], l5);
        if (nxt.is_zero()) throw std::runtime_error("zero pivot in the partial-round scaling");
        const F inv_n = inv(nxt);
        vs[j].resize(T - 1);
        ws[j].resize(T - 1);
        for (int i = 0; i < T - 1; i++) {
            vs[j][i] = mul(row0[j][i + 1], inv_n);
            ws[j][i] = mul(wcol[j][i], l5);
        }
        ks[j] = mul(kv(j), inv_n);
        lam = nxt;
    }
    const F lam5 = pow5(lam);

    put_canon(tbl, L::R2, R2);
    for (int i = 0; i < T; i++) put_v(tbl, L::IN_V + i, C(0, i));
    put_mont(tbl, L::S0, C(0, 0));
    for (int i = 0; i < T * T; i++) {
        put_mont(tbl, L::FULL_M + i, M[i]);
        put_mont(tbl, L::PRE_M + i, PRE[i]);
        put_mont(tbl, L::TAIL0_M + i, i % T == 0 ? mul(M[i], lam5) : M[i]);
    }
    for (int r = 0; r < 3; r++)
        for (int i = 0; i < T; i++) put_v(tbl, L::FULL_V + r * T + i, C(r + 1, i));
    put_v(tbl, L::PRE_V + 0, k[0]);     // remaining PRE_V entries stay zero
    for (int jp = 0; jp < L::N_PAIRS; jp++) {
        const int a = 2 * jp, b = 2 * jp + 1, base = L::PART + jp * L::PAIR_STRIDE;
        for (int i = 0; i < T - 1; i++) put_mont(tbl, base + L::P_VA + i, vs[a][i]);
        put_v(tbl, base + L::P_KA, ks[a]);
        for (int i = 0; i < T - 1; i++) put_mont(tbl, base + L::P_VB + i, vs[b][i]);
        F c = zero();                                          // c_B = v'_B . w'_A
        for (int i = 0; i < T - 1; i++) c = add(c, mul(vs[b][i], ws[a][i]));
        put_mont(tbl, base + L::P_CB, c);
        put_v(tbl, base + L::P_KB, ks[b]);
        for (int i = 0; i < T - 1; i++) {
            put_mont(tbl, base + L::P_W + 2 * i, ws[a][i]);
            put_mont(tbl, base + L::P_W + 2 * i + 1, ws[b][i]);
        }
    }
    for (int js = 0; js < L::N_SINGLES; js++) {
        const int j = 2 * L::N_PAIRS + js, base = L::SINGLES + js * L::SINGLE_STRIDE;
        for (int i = 0; i < T - 1; i++) put_mont(tbl, base + L::S_V + i, vs[j][i]);
        for (int i = 0; i < T - 1; i++) put_mont(tbl, base + L::S_W + i, ws[j][i]);
        put_v(tbl, base + L::S_K, ks[j]);
    }
    for (int j = 1; j < rp; j++) {
        F c = zero();
        for (int i = 0; i < T - 1; i++) c = add(c, mul(vs[j][i], ws[j - 1][i]));
        put_mont(tbl, L::COOP_C + j, c);
    }
    for (int i = 1; i < T; i++) put_mont(tbl, L::LAST_D + i - 1, D[i]);
    if constexpr (L::HR) {
        // history recurrence (Layout::HR; derive_hr in tests/opt_model.py)
        const int n = T - 1;
        auto dotv = [&](const std::vector<F>& x, const std::vector<F>& y) {
            F acc = zero();
            for (int i = 0; i < n; i++) acc = add(acc, mul(x[i], y[i]));
            return acc;
        };
        for (int c = 0; c < T; c++) put_mont(tbl, L::HR_PRE_M + c, PRE[c]);
        for (int j = 0; j < n; j++)
            for (int c = 0; c < T; c++) {
                F acc = zero();
                for (int x = 0; x < n; x++) acc = add(acc, mul(vs[j][x], PRE[(1 + x) * T + c]));
                put_mont(tbl, L::HR_PRE_M + (1 + j) * T + c, acc);
            }
        put_v(tbl, L::HR_PRE_V + 0, k[0]);
        for (int j = 0; j < n; j++) put_v(tbl, L::HR_PRE_V + 1 + j, ks[j]);
        for (int j = 1; j < n; j++)
            for (int i = j - 1; i >= 0; i--)                       // newest z first
                put_mont(tbl, L::HR_BOOT + j * (j - 1) / 2 + (j - 1 - i), dotv(vs[j], ws[i]));
        // row for a functional f (constant kf) of the passive state at round j, over
        // (u_j, .., u_{j-n+1}; z_{j-1}, .., z_{j-n})
        auto put_row = [&](int base, int j, const std::vector<F>& f, const F& kf) {
            Mat A(n * n);
            for (int i = 1; i <= n; i++)
                for (int x = 0; x < n; x++) A[(i - 1) * n + x] = vs[j - i][x];
            const Mat Ai = matinv(A, n);
            std::vector<F> g(n, zero());                           // f . A^-1
            for (int i = 0; i < n; i++)
                for (int x = 0; x < n; x++) g[i] = add(g[i], mul(f[x], Ai[x * n + i]));
            F cst = kf;
            for (int i = 1; i <= n; i++) {
                put_mont(tbl, base + i - 1, g[i - 1]);
                cst = sub(cst, mul(g[i - 1], ks[j - i]));
            }
            for (int l = 1; l <= n; l++) {
                F be = neg(g[l - 1]);
                for (int i = l; i <= n; i++) be = add(be, mul(g[i - 1], dotv(vs[j - i], ws[j - l])));
                put_mont(tbl, base + n + l - 1, be);
            }
            put_v(tbl, base + 2 * n, cst);
        };
        for (int j = n; j < rp; j++) put_row(L::HR_PART + (j - n) * L::HR_STRIDE, j, vs[j], ks[j]);
        for (int i = 0; i < n; i++) {
            std::vector<F> e(n, zero());
            e[i] = one();
            put_row(L::HR_EXIT + i * L::HR_STRIDE, rp, e, D[1 + i]);
        }
    }
    {
        // round 0 on unconverted inputs (Layout::R0_M): M R^6, (C_0[0])^5 / R^4, C_0 canonical.
        // to_mont multiplies by R, from_mont divides by it.
        for (int i = 0; i < T * T; i++) {
            F x = M[i];
            for (int e = 0; e < 6; e++) x = to_mont(x);
            to_limbs32(x, &tbl[(L::R0_M + i) * 8]);
        }
        F x0 = pow5(C(0, 0));
        for (int e = 0; e < 4; e++) x0 = from_mont(x0);
        to_limbs32(x0, &tbl[L::X0 * 8]);
        for (int i = 0; i < T; i++) put_canon(tbl, L::IN_C + i, C(0, i));
    }
    for (int r = 0; r < 3; r++)
        for (int i = 0; i < T; i++) put_v(tbl, L::TAIL_V + r * T + i, C(half + rp + r + 1, i));
    for (int j = 0; j < T; j++) {
        put_canon(tbl, L::OUT_ROW + j, M[j]);
        put_mont(tbl, L::OUT_ROW_MONT + j, M[j]);
    }
    return tbl;
}
}  // namespace

static std::vector<uint32_t> build_opt_uncached(int t) {
    switch (t) {
        case 2: return build_opt<2>();
        case 3: return build_opt<3>();
        case 4: return build_opt<4>();
        case 5: return build_opt<5>();
        case 6: return build_opt<6>();
        case 7: return build_opt<7>();
        case 8: return build_opt<8>();
        default: return {};
    }
}

// Derived once per process (a context per device asks for the same tables).
std::vector<uint32_t> build_opt_table(int t) {
    static std::map<int, std::vector<uint32_t>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(t);
    if (it == cache.end()) it = cache.emplace(t, build_opt_uncached(t)).first;
    return it->second;
}

std::vector<uint32_t> build_dense_table(int t) {
    const DenseParams& d = grain_params(t);
    std::vector<uint32_t> tbl((d.ark.size() + d.mds.size()) * 8);
    size_t k = 0;
    for (const F& x : d.ark) to_limbs32(to_mont(x), &tbl[8 * k++]);
    for (const F& x : d.mds) to_limbs32(to_mont(x), &tbl[8 * k++]);
    return tbl;
}

// pallet/src/poll/zeroes.rs:2 — 6769006970205099520508948723718471724660867171122235270773600567925038008762
const uint8_t BINARY_ZERO_LEAF_BE[32] = {
    0x0e, 0xf7, 0x1f, 0x46, 0xe1, 0x1a, 0x51, 0x3c, 0x59, 0x9e, 0xed, 0x9d, 0xd0, 0x35, 0x76, 0xc3,
    0x34, 0x39, 0xbc, 0xfb, 0x1c, 0xee, 0x15, 0x53, 0x16, 0xf9, 0x05, 0x41, 0xe4, 0x16, 0x49, 0xba};
// pallet/src/poll/zeroes.rs:38 — 8370432830353022751713833565135785980866757267633941821328460903436894336785
const uint8_t QUINARY_ZERO_LEAF_BE[32] = {
    0x12, 0x81, 0x7f, 0x41, 0x61, 0xf2, 0xf5, 0xde, 0xd3, 0x3f, 0x26, 0xc5, 0x57, 0x35, 0xa7, 0x7e,
    0x80, 0xe4, 0xf8, 0x97, 0x54, 0x83, 0xc8, 0xc2, 0x70, 0x47, 0x45, 0x12, 0x84, 0x17, 0xf7, 0x11};
// pallet/src/poll/zeroes.rs:73-79
const uint8_t EMPTY_BALLOT_ROOTS_BE[5][32] = {
    {0x23, 0x68, 0x7e, 0xc2, 0xcd, 0x45, 0x05, 0xa5, 0xb4, 0x3e, 0x69, 0xcb, 0xaf, 0x04, 0xdb, 0xa3,
     0x38, 0x4b, 0x19, 0xc5, 0x92, 0x37, 0xf9, 0x0f, 0x9e, 0x09, 0xed, 0x1c, 0xec, 0x66, 0x1b, 0xed},
    {0x00, 0x5e, 0x3d, 0xca, 0x15, 0xd7, 0x16, 0x9e, 0xc4, 0x24, 0xd8, 0xdf, 0x83, 0xa9, 0xe7, 0xb4,
     0xa4, 0x3e, 0xbd, 0xe4, 0xf2, 0x0b, 0xde, 0xcc, 0x04, 0x89, 0x77, 0x79, 0xbc, 0xd1, 0x57, 0x3b},
    {0x16, 0x3c, 0x79, 0xcb, 0x43, 0x33, 0x05, 0x00, 0xfc, 0x1f, 0x32, 0x26, 0x99, 0xdf, 0xb8, 0xa5,
     0x0d, 0xe8, 0x4f, 0x6f, 0xeb, 0xc1, 0x3c, 0x15, 0x03, 0xb2, 0x14, 0xec, 0xd9, 0x6d, 0x0c, 0x01},
    {0x0a, 0xd8, 0x08, 0xdc, 0xd3, 0xf8, 0xd3, 0xb2, 0xea, 0xe3, 0x85, 0x08, 0x3b, 0x78, 0x08, 0x2e,
     0x65, 0xe1, 0x56, 0x9b, 0x3b, 0x30, 0xa8, 0x45, 0x98, 0x8f, 0x14, 0x59, 0x4a, 0x2b, 0x28, 0x95},
    {0x29, 0x54, 0x76, 0xdc, 0x2c, 0x5c, 0x66, 0x1b, 0x50, 0x71, 0xdc, 0x89, 0x4c, 0x75, 0x60, 0xa0,
     0xab, 0x24, 0x08, 0x61, 0x3a, 0xcd, 0x6e, 0x2c, 0x79, 0x4c, 0xde, 0x85, 0x28, 0xe2, 0x29, 0x09}};

}  // namespace host
}  // namespace inf
