// developer tool: how fast do T host threads READ 3 GB (AVX-512 loads, nothing else) from (a) malloc'ed memory, (b) the same
// without transparent huge pages, (c) cudaHostAlloc memory, (d) 2 MB-aligned memory with MADV_HUGEPAGE + cudaHostRegister —
// and how fast does the copy engine take (c) and (d)?  The packed upload reads every byte of the pinned arenas once.
// nvcc -O3 -Xcompiler -mavx512f,-pthread tools/host_read_bw.cu -o /tmp/hrb2
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double read_bw(const uint8_t* src, size_t bytes, int T, int pf) {
    double best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        std::vector<uint64_t> sink(T * 16);
        for (int t = 0; t < T; t++) th.emplace_back([&, t] {
            size_t a = (bytes / 256 * t / T) * 256, b = (bytes / 256 * (t + 1) / T) * 256;
            __m512i acc = _mm512_setzero_si512();
            for (size_t i = a; i < b; i += 256) {
                if (pf) { _mm_prefetch((const char*)src + i + pf, _MM_HINT_T0); _mm_prefetch((const char*)src + i + pf + 128, _MM_HINT_T0); }
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
    return bytes / 1e9 / best;
}
static double h2d_bw(const void* src, void* dst, size_t bytes) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a); cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return bytes / 1e6 / best;
}
int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? atol(argv[1]) : 3000) * 1000000ull / (2 << 20) * (2 << 20);
    void* dev; cudaMalloc(&dev, bytes);
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r"); char buf[128] = {0}; if (f) { fgets(buf, 127, f); fclose(f); } printf("THP: %s", buf);
    for (int kind = 0; kind < 4; kind++) {
        uint8_t* p = nullptr; const char* name = "";
        if (kind == 0) { p = (uint8_t*)aligned_alloc(2 << 20, bytes); name = "malloc (THP default)"; }
        if (kind == 1) { p = (uint8_t*)aligned_alloc(2 << 20, bytes); madvise(p, bytes, MADV_NOHUGEPAGE); name = "malloc, MADV_NOHUGEPAGE"; }
        if (kind == 2) { cudaHostAlloc((void**)&p, bytes, cudaHostAllocDefault); name = "cudaHostAlloc"; }
        if (kind == 3) { p = (uint8_t*)aligned_alloc(2 << 20, bytes); madvise(p, bytes, MADV_HUGEPAGE); name = "MADV_HUGEPAGE + cudaHostRegister"; }
        auto t0 = std::chrono::steady_clock::now();
        memset(p, 'A', bytes);
        if (kind == 3) { cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault); if (e != cudaSuccess) printf("register failed: %s\n", cudaGetErrorString(e)); }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("%-36s touch%s %.2f s;", name, kind == 3 ? " + register" : "", s);
        for (int T : {1, 12}) for (int pf : {0, 8192}) printf("  T%d pf%d: %.1f GB/s", T, pf, read_bw(p, bytes, T, pf));
        if (kind >= 2) printf("  H2D %.1f GB/s", h2d_bw(p, dev, bytes));
        printf("\n"); fflush(stdout);
        if (kind == 2) cudaFreeHost(p); else { if (kind == 3) cudaHostUnregister(p); free(p); }
    }
    return 0;
}
