// developer tool: how fast can the host turn ASCII reads into 2-bit planes (AVX-512BW, N threads) compared with memcpy?  g++ -O3 -mavx512f -mavx512bw -mavx512vl -pthread
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static inline void pack_read(const uint8_t* s, int len, uint32_t* lo, uint32_t* hi, uint32_t* bad) {
    const __m512i b4 = _mm512_set1_epi8(4), b2 = _mm512_set1_epi8(2), m7 = _mm512_set1_epi8(7);
    // expected letter by low 3 bits: 1 A, 3 C, 7 G, 4 T
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1));
    int w = 0; uint64_t anybad = 0;
    for (int p = 0; p < len; p += 64, w += 2) {
        const int rem = len - p;
        const __mmask64 km = rem >= 64 ? ~0ull : ((1ull << rem) - 1ull);
        const __m512i x = _mm512_maskz_loadu_epi8(km, s + p);
        const uint64_t l = _mm512_test_epi8_mask(x, b4), h = _mm512_test_epi8_mask(x, b2);
        const __m512i e = _mm512_shuffle_epi8(lut, _mm512_and_si512(x, m7));
        const uint64_t v = _mm512_mask_cmpeq_epi8_mask(km, x, e);
        anybad |= v ^ km;
        lo[w] = (uint32_t)(l & v); hi[w] = (uint32_t)(h & v);
        if (rem > 32) { lo[w + 1] = (uint32_t)((l & v) >> 32); hi[w + 1] = (uint32_t)((h & v) >> 32); }
    }
    *bad |= anybad != 0;
}
int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 20000000; const int L = 150; const int T = argc > 2 ? atoi(argv[2]) : 16;
    std::vector<uint8_t> src((size_t)n * L + 64);
    for (size_t i = 0; i < src.size(); i++) src[i] = "ACGT"[(i * 2654435761u >> 13) & 3];
    const int nw = (L + 31) / 32;
    std::vector<uint32_t> dst((size_t)n * nw * 2 + 64);
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th; std::vector<uint32_t> bad(T * 16, 0);
        for (int t = 0; t < T; t++) th.emplace_back([&, t] {
            const long a = n * t / T, b = n * (t + 1) / T;
            for (long i = a; i < b; i++) pack_read(src.data() + (size_t)i * L, L, dst.data() + (size_t)i * nw * 2, dst.data() + (size_t)i * nw * 2 + nw, &bad[t * 16]);
        });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.1f ms for %.2f GB -> %.1f GB/s in, %.1f M reads/s\n", T, s * 1e3, n * (double)L / 1e9, n * (double)L / 1e9 / s, n / s / 1e6);
    }
    // plain memcpy for comparison
    std::vector<uint8_t> cp(src.size());
    for (int rep = 0; rep < 2; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t] { size_t a = src.size() * t / T, b = src.size() * (t + 1) / T; memcpy(cp.data() + a, src.data() + a, b - a); });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("memcpy %d threads: %.1f GB/s\n", T, src.size() / 1e9 / s);
    }
    return 0;
}
