// Host build of the device arithmetic (fr.cuh / poseidon.cuh with the PTX
// blocks replaced by their C bodies) behind a tiny C interface, so that the
// CPU test-suite can exercise the exact kernel logic where there is no GPU.
// TEST SUPPORT: built as libinfimum_hostemu.so, never loaded by the product.
#define INF_HOST_CHECKS 1
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "coop.cuh"
#include "host_params.h"
#include "poseidon.cuh"

using namespace inf;

template <int T>
static void hash_t(const uint8_t* in, const uint8_t* tag, uint8_t* out, int le, const uint32_t* tbl) {
    uint32_t iw[T - 1][8], ow[8], tw[8];
    memcpy(iw, in, (T - 1) * 32);
    if (tag) memcpy(tw, tag, 32);
    if (le) hash_words<T, true>(ow, iw, tag ? tw : nullptr, tbl);
    else    hash_words<T, false>(ow, iw, tag ? tw : nullptr, tbl);
    memcpy(out, ow, 32);
}

// ---- the warp-cooperative schedule (coop.cuh) with one host thread per role -------------
// Barriers follow the PTX named-barrier rules the device Bus relies on: a generation
// completes when the expected number of roles has entered (arrive = enter without
// waiting, wait = enter and block); a role entering the same barrier twice in one
// generation is a protocol error and is counted.
namespace {
struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int count = 0;
    unsigned entered = 0;            // bit per role, this generation
    unsigned long long gen = 0;
    std::atomic<int>* errors = nullptr;
    void enter(int role, int expected, bool block) {
        std::unique_lock<std::mutex> lk(m);
        if (entered & (1u << role)) ++*errors;
        entered |= 1u << role;
        if (++count == expected) {
            count = 0;
            entered = 0;
            gen++;
            cv.notify_all();
        } else if (block) {
            const unsigned long long g = gen;
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

template <int T>
struct HostBus {
    typedef uint32_t* Slot;
    uint32_t xs[2][T][8], zs[3][8], vs[2][8], ps[3][T][8];
    HostBarrier blk, zb[3], vb[2], pb[3], pro;
    std::atomic<int> errors{0};
    unsigned jitter = 0;             // roles (bit mask) that dawdle before every barrier
    HostBus() {
        for (HostBarrier* b : {&blk, &zb[0], &zb[1], &zb[2], &vb[0], &vb[1], &pb[0], &pb[1], &pb[2], &pro}) b->errors = &errors;
    }
    struct View {                    // what one role's thread sees
        HostBus& b;
        int role;
        Slot x(int buf, int i) const { return b.xs[buf][i]; }
        Slot z(int par) const { return b.zs[par]; }
        Slot v(int par) const { return b.vs[par]; }
        Slot p(int par, int i) const { return b.ps[par][i]; }
        void put(Slot d, const uint32_t (&val)[8]) const { memcpy(d, val, 32); }
        void get(uint32_t (&val)[8], Slot s) const { memcpy(val, s, 32); }
        void dawdle() const {
            if (b.jitter & (1u << role)) std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
        void block() const { dawdle(); b.blk.enter(role, T + 1, true); }
        void z_arrive(int par) const { dawdle(); b.zb[par].enter(role, T + 1, false); }
        void z_wait(int par) const { dawdle(); b.zb[par].enter(role, T + 1, true); }
        void v_arrive(int par) const { dawdle(); b.vb[par].enter(role, 2, false); }
        void v_wait(int par) const { dawdle(); b.vb[par].enter(role, 2, true); }
        void p_arrive(int par) const { dawdle(); b.pb[par].enter(role, T, false); }
        void p_wait(int par) const { dawdle(); b.pb[par].enter(role, T, true); }
        void pro_arrive() const { dawdle(); b.pro.enter(role, T, false); }
        void pro_wait() const { dawdle(); b.pro.enter(role, T, true); }
    };
};

template <int T>
int coop_hash_t(const uint8_t* in, uint8_t* out, unsigned jitter, const uint32_t* tbl) {
    using L = Layout<T>;
    HostBus<T> bus;
    bus.jitter = jitter;
    uint32_t result[8] = {};
    std::thread th[T + 1];
    for (int role = 0; role <= T; role++)
        th[role] = std::thread([&, role] {
            uint32_t s[8] = {}, h[8];
            if (role == 0) {
                memcpy(s, tbl + L::S0 * 8, 32);
            } else if (role < T) {
                uint32_t wd[8], raw[8];
                memcpy(wd, in + 32 * (role - 1), 32);
                words_to_limbs<false>(raw, wd);
                absorb<T>(s, raw, role, tbl);
            }
            typename HostBus<T>::View view{bus, role};
            coop_hash<T>(h, s, role, view, tbl);
            if (role == 0) memcpy(result, h, 32);
        });
    for (auto& t : th) t.join();
    uint32_t ow[8];
    limbs_to_words<false>(ow, result);
    memcpy(out, ow, 32);
    return bus.errors.load();
}
}  // namespace

extern "C" {

// One hash through the warp-cooperative schedule, one thread per role.  Returns the
// number of barrier-protocol errors (0 expected), or -1 for an unsupported width.
int hostemu_coop_hash(int t, const uint8_t* in, uint8_t* out, unsigned jitter) {
    static std::vector<uint32_t> tables[14];
    static std::mutex mu;
    if (t < 2 || t > 8) return -1;
    {
        std::lock_guard<std::mutex> g(mu);
        if (tables[t].empty()) tables[t] = host::build_opt_table(t);
    }
    const uint32_t* tbl = tables[t].data();
    switch (t) {
        case 2: return coop_hash_t<2>(in, out, jitter, tbl);
        case 3: return coop_hash_t<3>(in, out, jitter, tbl);
        case 4: return coop_hash_t<4>(in, out, jitter, tbl);
        case 5: return coop_hash_t<5>(in, out, jitter, tbl);
        case 6: return coop_hash_t<6>(in, out, jitter, tbl);
        case 7: return coop_hash_t<7>(in, out, jitter, tbl);
        case 8: return coop_hash_t<8>(in, out, jitter, tbl);
    }
    return -1;
}

unsigned long long hostemu_overflow_count() { return host_overflow_count; }

void hostemu_mont_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    uint32_t t[8];
    mont_mul(t, a, b);
    memcpy(r, t, 32);
}
void hostemu_mont_sqr(const uint32_t* a, uint32_t* r) {
    uint32_t t[8];
    mont_sqr(t, a);
    memcpy(r, t, 32);
}
void hostemu_mont_mul_add(const uint32_t* a, const uint32_t* b, const uint32_t* v, uint32_t* r) {
    uint32_t t[8];
    mont_mul_add(t, a, b, v);
    memcpy(r, t, 32);
}
void hostemu_redc(const uint32_t* x, uint32_t* r) {
    uint32_t t[8];
    mont_redc(t, x);
    memcpy(r, t, 32);
}
// n-term lazy dot, n in 1..10: a = n x 8 limbs, b = n x 8 limbs, v = 8 limbs or NULL
int hostemu_dot(int n, const uint32_t* a, const uint32_t* b, const uint32_t* v, uint32_t* r) {
    uint32_t t[8];
    switch (n) {
        case 1: dot<1, 8>(t, a, b, v); break;
        case 2: dot<2, 8>(t, a, b, v); break;
        case 3: dot<3, 8>(t, a, b, v); break;
        case 4: dot<4, 8>(t, a, b, v); break;
        case 5: dot<5, 8>(t, a, b, v); break;
        case 6: dot<6, 8>(t, a, b, v); break;
        case 7: dot<7, 8>(t, a, b, v); break;
        case 8: dot<8, 8>(t, a, b, v); break;
        case 9: dot<9, 8>(t, a, b, v); break;
        case 10: dot<10, 8>(t, a, b, v); break;     // the widest row of the history recurrence (t = 6)
        default: return -1;
    }
    memcpy(r, t, 32);
    return 0;
}
void hostemu_csub2p(uint32_t* x) {
    uint32_t t[8];
    memcpy(t, x, 32);
    csub2p(t);
    memcpy(x, t, 32);
}
void hostemu_csub4p(uint32_t* x) {
    uint32_t t[8];
    memcpy(t, x, 32);
    csub4p(t);
    memcpy(x, t, 32);
}
// Range steps of a history-recurrence row of n passive elements (hr_range<N>, poseidon.cuh).
void hostemu_hr_range(int n, uint32_t* x) {
    uint32_t t[8];
    memcpy(t, x, 32);
    if (n >= 5) hr_range<5>(t); else hr_range<1>(t);
    memcpy(x, t, 32);
}
void hostemu_csub_p_exact(uint32_t* x) {
    uint32_t t[8];
    memcpy(t, x, 32);
    csub_p_exact(t);
    memcpy(x, t, 32);
}

// One hash through the optimised schedule.  width t in 2..8.  Returns 0, or -1.
int hostemu_hash(int t, const uint8_t* in, const uint8_t* tag, uint8_t* out, int le) {
    static std::vector<uint32_t> tables[14];
    if (t < 2 || t > 8) return -1;
    if (tables[t].empty()) tables[t] = host::build_opt_table(t);
    const uint32_t* tbl = tables[t].data();
    switch (t) {
        case 2: hash_t<2>(in, tag, out, le, tbl); break;
        case 3: hash_t<3>(in, tag, out, le, tbl); break;
        case 4: hash_t<4>(in, tag, out, le, tbl); break;
        case 5: hash_t<5>(in, tag, out, le, tbl); break;
        case 6: hash_t<6>(in, tag, out, le, tbl); break;
        case 7: hash_t<7>(in, tag, out, le, tbl); break;
        case 8: hash_t<8>(in, tag, out, le, tbl); break;
    }
    return 0;
}

// Table access for cross-checks against the Python derivation.
int hostemu_opt_table_words(int t) { return (int)host::build_opt_table(t).size(); }
int hostemu_opt_table(int t, uint32_t* out) {
    std::vector<uint32_t> v = host::build_opt_table(t);
    memcpy(out, v.data(), v.size() * 4);
    return (int)v.size();
}
int hostemu_dense_params(int t, uint32_t* out /* canonical limbs: ark then mds */) {
    const host::DenseParams& d = host::grain_params(t);
    size_t k = 0;
    for (const host::F& x : d.ark) host::to_limbs32(x, out + 8 * k++);
    for (const host::F& x : d.mds) host::to_limbs32(x, out + 8 * k++);
    return (int)k;
}
}
