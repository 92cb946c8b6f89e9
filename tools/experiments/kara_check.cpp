// Host check of the Karatsuba experiment against the library's own dot product.
//   g++ -O2 -std=c++17 -DINF_HOST_CHECKS -I infimum_b200/csrc -I tools/experiments tools/experiments/kara_check.cpp -o tools/_bin/kara_check
#include <cstdio>
#include <random>
#include "poseidon.cuh"
#include "kara.cuh"
using namespace inf;
template <int N> int run(std::mt19937& g) {
    int bad = 0;
    for (int it = 0; it < 20000; it++) {
        uint32_t a[N * 8], b[N * 8], bs[N * 5], r0[8], r1[8];
        for (int k = 0; k < N * 8; k++) { a[k] = g(); b[k] = g(); }
        for (int j = 0; j < N; j++) {
            a[8 * j + 7] &= (it & 1) ? 0x7fffffffu : 0x3fffffffu;
            b[8 * j + 7] &= 0x3fffffffu; if (b[8 * j + 7] > INF_P7 - 1) b[8 * j + 7] = INF_P7 - 1;
            if (it % 7 == 0) for (int k = 0; k < 4; k++) a[8 * j + k] = a[8 * j + 4 + k] = 0xffffffffu >> (k == 3 && true ? 1 : 0);
            if (it % 11 == 0) for (int k = 0; k < 4; k++) { b[8 * j + k] = 0xffffffffu; }
        }
        half_sums(bs, b, N);
        dot<N, 8, false>(r0, a, b, nullptr);
        dot_kara<N, 8>(r1, a, b, bs);
        for (int k = 0; k < 8; k++) if (r0[k] != r1[k]) { bad++; break; }
    }
    return bad;
}
int main() {
    std::mt19937 g(7);
    int bad = run<1>(g) + run<2>(g) + run<3>(g) + run<4>(g);
    printf("mismatches %d, overflow flags %llu\n", bad, host_overflow_count);
    return bad != 0;
}
