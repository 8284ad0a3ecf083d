// developer tool: how fast can T host threads READ memory (AVX-512 loads, nothing else)?  The packed upload reads every byte of the
// arenas once, so this is its ceiling.  g++ -O3 -mavx512f -pthread tools/host_read_bw.cpp -o /tmp/host_read_bw
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? atol(argv[1]) : 3000) * 1000000ull;
    uint8_t* src = (uint8_t*)aligned_alloc(4096, bytes);
    memset(src, 'A', bytes);
    for (int T : {1, 4, 8, 12, 14, 16}) {
        for (int pf : {0, 8192}) {
            double best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                std::vector<uint64_t> sink(T * 16);
                for (int t = 0; t < T; t++) th.emplace_back([&, t] {
                    size_t a = (bytes / 64 * t / T) * 64, b = (bytes / 64 * (t + 1) / T) * 64;
                    __m512i acc = _mm512_setzero_si512();
                    for (size_t i = a; i < b; i += 256) {
                        if (pf) { _mm_prefetch((const char*)src + i + pf, _MM_HINT_T0); _mm_prefetch((const char*)src + i + pf + 64, _MM_HINT_T0);
                                  _mm_prefetch((const char*)src + i + pf + 128, _MM_HINT_T0); _mm_prefetch((const char*)src + i + pf + 192, _MM_HINT_T0); }
                        acc = _mm512_or_si512(acc, _mm512_load_si512((const void*)(src + i)));
                        acc = _mm512_or_si512(acc, _mm512_load_si512((const void*)(src + i + 64)));
                        acc = _mm512_or_si512(acc, _mm512_load_si512((const void*)(src + i + 128)));
                        acc = _mm512_or_si512(acc, _mm512_load_si512((const void*)(src + i + 192)));
                    }
                    sink[t * 16] = _mm512_reduce_or_epi64(acc);
                });
                for (auto& x : th) x.join();
                double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (s < best) best = s;
            }
            printf("read %2d threads prefetch %5d: %.1f GB/s\n", T, pf, bytes / 1e9 / best);
        }
    }
    return 0;
}
