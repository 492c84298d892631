// sincos_ref(x) == (sin_ref(x), cos_ref(x)) bit for bit, for every `stride`-th float bit pattern (stride 1 = all 2^32) and for
// the neighbourhoods of every range boundary of the two functions.  Test infrastructure (built by tests/test_trig.py).
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../resolve2d_b200/csrc/r2d_math.cuh"
using namespace r2d;

static bool same(float x) {
    float s, c;
    sincos_ref(x, &s, &c);
    const float s0 = sin_ref(x), c0 = cos_ref(x);
    if (s != s || s0 != s0 || c != c || c0 != c0) return (s != s) == (s0 != s0) && (c != c) == (c0 != c0);
    return f2u(s) == f2u(s0) && f2u(c) == f2u(c0);
}

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const int T = (int)std::thread::hardware_concurrency() > 0 ? (int)std::thread::hardware_concurrency() : 4;
    std::atomic<uint64_t> bad{0}, n{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([=, &bad, &n]() {
            uint64_t k = 0;
            for (uint64_t b = (uint64_t)t * stride; b < (1ull << 32); b += (uint64_t)T * stride, ++k)
                if (!same(u2f((uint32_t)b)) && bad++ < 10) printf("MISMATCH at bits %08x\n", (uint32_t)b);
            n += k;
        });
    for (auto& x : th) x.join();
    const uint32_t edges[] = {0x39800000u, 0x3f490fdau, 0x4016cbe3u, 0x407b53d1u, 0x40afeddfu, 0x40e231d5u, 0x4dc90fdbu, 0x7f800000u, 0u};
    for (uint32_t e : edges)
        for (int dlt = -4096; dlt <= 4096; ++dlt)
            for (uint32_t sgn : {0u, 0x80000000u}) {
                const uint32_t b = (uint32_t)((int64_t)e + dlt) | sgn;
                n += 1;
                if (!same(u2f(b)) && bad++ < 10) printf("MISMATCH at bits %08x\n", b);
            }
    printf("checked %llu arguments, mismatches: %llu\n", (unsigned long long)n.load(), (unsigned long long)bad.load());
    return bad ? 1 : 0;
}
