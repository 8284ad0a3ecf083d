/*
 * replay.c — plain-C driver over the C ABI (include/genefuse_gpu.h), the way the Rust shim would drive it:
 *   pack of 1000 pairs (src/core/common.rs:23) -> [gf_stream_push | gf_batch + gf_map_pairs] -> records.
 * Proves that the boundary is usable without Python / torch, and measures what a call costs at small batch sizes
 * (VERDICT r1 #8): one gf_map_pairs call per S pairs (S = 1 k / 64 k / 1 M) against the batched shim fed with 1000-pair packs.
 *
 * usage: replay <dump file written by tests/test_c_driver.py> [max_pairs]
 * dump:  u32 n_genes, then per gene {u32 len, u8 reversed, bytes}; u64 n_pairs, u32 L; seq1, qual1, seq2, qual2 (n * L each)
 * prints one JSON line per mode.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "genefuse_gpu.h"

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static int cmp_rec(const void* a, const void* b) {
    const gf_match *x = (const gf_match*)a, *y = (const gf_match*)b;
    if (x->pair_idx != y->pair_idx) return x->pair_idx < y->pair_idx ? -1 : 1;
    return (int)x->source - (int)y->source;
}
static uint64_t checksum(gf_match* r, uint64_t n) {
    qsort(r, n, sizeof(gf_match), cmp_rec);
    uint64_t h = 1469598103934665603ull;
    const unsigned char* p = (const unsigned char*)r;
    for (uint64_t i = 0; i < n * sizeof(gf_match); i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
#define CHECK(rc, what) do { if ((rc) != GF_OK) { fprintf(stderr, "%s failed: %d %s\n", what, (rc), gf_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: replay dump [max_pairs]\n"); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("dump"); return 1; }
    uint32_t n_genes = 0;
    if (fread(&n_genes, 4, 1, f) != 1) return 1;
    gf_gene_span* genes = (gf_gene_span*)calloc(n_genes ? n_genes : 1, sizeof(gf_gene_span));
    for (uint32_t g = 0; g < n_genes; g++) {
        uint32_t len; uint8_t rev;
        if (fread(&len, 4, 1, f) != 1 || fread(&rev, 1, 1, f) != 1) return 1;
        uint8_t* s = (uint8_t*)malloc(len ? len : 1);
        if (len && fread(s, 1, len, f) != len) return 1;
        genes[g].seq = s; genes[g].len = len; genes[g].reversed = rev;
    }
    uint64_t n = 0; uint32_t L = 0;
    if (fread(&n, 8, 1, f) != 1 || fread(&L, 4, 1, f) != 1) return 1;
    uint8_t* arena[4];
    for (int k = 0; k < 4; k++) {
        arena[k] = (uint8_t*)malloc(n * L);
        if (fread(arena[k], 1, n * L, f) != n * L) return 1;
    }
    fclose(f);
    if (argc > 2 && (uint64_t)atoll(argv[2]) < n) n = (uint64_t)atoll(argv[2]);

    gf_index* idx = NULL;
    double t0 = now_s();
    CHECK(gf_index_create(genes, n_genes, NULL, 0, &idx), "gf_index_create");
    fprintf(stderr, "index over %u genes in %.3f s (incl. CUDA init)\n", n_genes, now_s() - t0);

    uint64_t* off = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i <= n; i++) off[i] = i * L;
    const uint64_t cap = 2 * n + 16;
    gf_match* rec = (gf_match*)malloc(cap * sizeof(gf_match));
    gf_match* tmp = (gf_match*)malloc(cap * sizeof(gf_match));
    const uint64_t sizes[3] = {1000, 65536, 1048576};

    /* (a) one gf_map_pairs call per S pairs, straight from the (pageable) arenas */
    for (int si = 0; si < 3; si++) {
        const uint64_t S = sizes[si];
        uint64_t total = 0, calls = 0;
        for (int rep = 0; rep < 2; rep++) { /* rep 0 warms the handle's buffers up */
            total = 0; calls = 0;
            t0 = now_s();
            for (uint64_t lo = 0; lo < n; lo += S) {
                const uint64_t hi = lo + S < n ? lo + S : n;
                gf_batch b;
                memset(&b, 0, sizeof(b));
                b.n = hi - lo;
                b.seq1 = arena[0]; b.qual1 = arena[1]; b.off1 = off + lo;
                b.seq2 = arena[2]; b.qual2 = arena[3]; b.off2 = off + lo;
                b.bytes1 = b.bytes2 = (hi - lo) * L;
                b.max_len = L;
                uint64_t got = 0;
                CHECK(gf_map_pairs(idx, &b, tmp, cap, &got), "gf_map_pairs");
                for (uint64_t i = 0; i < got; i++) { tmp[i].pair_idx += lo; rec[total + i] = tmp[i]; }
                total += got;
                calls++;
            }
        }
        const double dt = now_s() - t0;
        printf("{\"mode\": \"gf_map_pairs per call\", \"call_pairs\": %llu, \"calls\": %llu, \"pairs\": %llu, \"seconds\": %.6f, "
               "\"pairs_per_s\": %.1f, \"us_per_call\": %.2f, \"records\": %llu, \"checksum\": \"%016llx\"}\n",
               (unsigned long long)S, (unsigned long long)calls, (unsigned long long)n, dt, n / dt, 1e6 * dt / calls,
               (unsigned long long)total, (unsigned long long)checksum(rec, total));
        fflush(stdout);
    }

    /* (b) the batched shim: packs of 1000 pairs as pointer arrays (what a ReadPairPack holds), batches of S pairs */
    const uint8_t** p[4];
    uint32_t* len = (uint32_t*)malloc(1000 * sizeof(uint32_t));
    for (int k = 0; k < 4; k++) p[k] = (const uint8_t**)malloc(1000 * sizeof(uint8_t*));
    for (int i = 0; i < 1000; i++) len[i] = L;
    for (int si = 0; si < 3; si++) {
        const uint64_t S = sizes[si];
        uint64_t total = 0, calls = 0;
        for (int rep = 0; rep < 2; rep++) {
            gf_stream* st = NULL;
            CHECK(gf_stream_create(idx, 1, S, &st), "gf_stream_create");
            t0 = now_s();
            for (uint64_t lo = 0; lo < n; lo += 1000) {
                const uint64_t cnt = lo + 1000 < n ? 1000 : n - lo;
                for (uint64_t i = 0; i < cnt; i++)
                    for (int k = 0; k < 4; k++) p[k][i] = arena[k] + (lo + i) * L;
                CHECK(gf_stream_push(st, lo, cnt, p[0], p[1], len, p[2], p[3], len), "gf_stream_push");
            }
            CHECK(gf_stream_flush(st), "gf_stream_flush");
            CHECK(gf_stream_take(st, rec, cap, &total), "gf_stream_take");
            uint64_t pushed = 0;
            gf_stream_get_counts(st, &pushed, &calls);
            gf_stream_destroy(st);
        }
        const double dt = now_s() - t0;
        printf("{\"mode\": \"gf_stream (1000-pair packs)\", \"call_pairs\": %llu, \"calls\": %llu, \"pairs\": %llu, \"seconds\": %.6f, "
               "\"pairs_per_s\": %.1f, \"us_per_call\": %.2f, \"records\": %llu, \"checksum\": \"%016llx\"}\n",
               (unsigned long long)S, (unsigned long long)calls, (unsigned long long)n, dt, n / dt, 1e6 * dt / calls,
               (unsigned long long)total, (unsigned long long)checksum(rec, total));
        fflush(stdout);
    }
    gf_index_destroy(idx);
    return 0;
}
