/*
 * gf_synth.cpp — synthetic workload generator (host, multi-threaded).  Bench/test utility,
 * not part of the matching path.  Implements the read model of SURVEY.md §8(d): pair i is a
 * pure function of (seed, i) (counter-based), so any shard can be produced on any rank.
 *
 *   fragment length  clip(round(N(1.8 L, 0.3 L)), L, 4 L)
 *   source           p_target: uniform over gene spans (either strand)
 *                    p_fusion: fragment spanning one of the planted breakpoints
 *                    rest:     i.i.d. random sequence (off target)
 *   R1 = frag[..L], R2 = rc(frag)[..L]; substitutions sub_rate/base, N n_rate/base
 *   qualities        'E' 0.90, 'A' 0.07, '/' 0.03
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {
struct Rng { /* splitmix64 stream keyed by (seed, counter) */
    uint64_t s;
    static uint64_t mix(uint64_t z) {
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t ctr) { s = mix(seed * 0x9E3779B97F4A7C15ULL + mix(ctr + 0x632BE59BD9B4E019ULL)); }
    uint64_t next() { s += 0x9E3779B97F4A7C15ULL; return mix(s); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(uni() * (double)n); }
};
inline uint8_t comp(uint8_t b) {
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}
const char BASES[4] = {'A', 'C', 'G', 'T'};
}  // namespace

extern "C" {

typedef struct gfs_fusion {
    int32_t gene_a, pos_a, strand_a; /* last base of the left part (strand +) / first (strand -) */
    int32_t gene_b, pos_b, strand_b;
} gfs_fusion;

typedef struct gfs_config {
    uint64_t seed;
    int32_t read_len;
    double p_target, p_fusion, sub_rate, n_rate;
    int32_t n_genes;
    const uint8_t* gene_arena;    /* concatenated gene bytes */
    const uint64_t* gene_off;     /* n_genes + 1 */
    int32_t n_fusions;
    const gfs_fusion* fusions;
} gfs_config;

/* i.i.d. uniform ACGT bytes, counter-based per 64-byte block */
void gfs_random_bases(uint64_t seed, uint8_t* out, uint64_t n) {
    for (uint64_t blk = 0; blk * 64 < n; blk++) {
        Rng r(seed, blk);
        for (uint64_t w = 0; w < 2; w++) {
            uint64_t x = r.next();
            for (int k = 0; k < 32; k++) {
                uint64_t i = blk * 64 + w * 32 + k;
                if (i < n) out[i] = BASES[(x >> (2 * k)) & 3];
            }
        }
    }
}

/* Generate pairs [first, first+count) into fixed-stride arenas (stride = read_len). */
void gfs_generate_pairs(const gfs_config* cfg, uint64_t first, uint64_t count, uint8_t* seq1, uint8_t* qual1,
                        uint8_t* seq2, uint8_t* qual2, uint8_t* kind_out, int threads) {
    const int L = cfg->read_len;
    uint64_t total_gene = cfg->gene_off[cfg->n_genes];
    auto work = [&](uint64_t lo, uint64_t hi) {
        std::vector<uint8_t> frag((size_t)4 * L + 8);
        for (uint64_t p = lo; p < hi; p++) {
            Rng r(cfg->seed, first + p);
            double u1 = r.uni(), u2 = r.uni();
            double z = std::sqrt(-2.0 * std::log(u1 + 1e-300)) * std::cos(6.283185307179586 * u2);
            long f = std::lround(1.8 * L + 0.3 * L * z);
            if (f < L) f = L;
            if (f > 4 * L) f = 4 * L;
            double us = r.uni();
            uint8_t kind = 1; /* 0 target, 1 off-target, 2 fusion */
            if (us < cfg->p_target && total_gene > 0) {
                /* uniform over gene bases; retry a few times until the fragment fits */
                for (int attempt = 0; attempt < 8 && kind != 0; attempt++) {
                    uint64_t g = r.below(total_gene);
                    int lo_g = 0, hi_g = cfg->n_genes;
                    while (hi_g - lo_g > 1) {
                        int mid = (lo_g + hi_g) / 2;
                        if (cfg->gene_off[mid] <= g) lo_g = mid; else hi_g = mid;
                    }
                    uint64_t glen = cfg->gene_off[lo_g + 1] - cfg->gene_off[lo_g];
                    uint64_t st = g - cfg->gene_off[lo_g];
                    if (st + (uint64_t)f <= glen) {
                        memcpy(frag.data(), cfg->gene_arena + cfg->gene_off[lo_g] + st, (size_t)f);
                        kind = 0;
                    }
                }
            } else if (us < cfg->p_target + cfg->p_fusion && cfg->n_fusions > 0) {
                const gfs_fusion& fu = cfg->fusions[r.below((uint64_t)cfg->n_fusions)];
                /* 10 distinct split points per fusion (duplicates AND unique reads exist) */
                long x = 40 + (long)r.below(10) * ((f - 80) / 10 + 1);
                if (x > f - 40) x = f - 40;
                if (x < 1) x = f / 2;
                long y = f - x;
                const uint8_t* ga = cfg->gene_arena + cfg->gene_off[fu.gene_a];
                const uint8_t* gb = cfg->gene_arena + cfg->gene_off[fu.gene_b];
                long la = (long)(cfg->gene_off[fu.gene_a + 1] - cfg->gene_off[fu.gene_a]);
                long lb = (long)(cfg->gene_off[fu.gene_b + 1] - cfg->gene_off[fu.gene_b]);
                bool ok = true;
                if (fu.strand_a > 0) { if (fu.pos_a - x + 1 < 0) ok = false; }
                else { if (fu.pos_a + x > la) ok = false; }
                if (fu.strand_b > 0) { if (fu.pos_b + y > lb) ok = false; }
                else { if (fu.pos_b - y + 1 < 0) ok = false; }
                if (ok) {
                    for (long k = 0; k < x; k++)
                        frag[(size_t)k] = fu.strand_a > 0 ? ga[fu.pos_a - x + 1 + k] : comp(ga[fu.pos_a + x - 1 - k]);
                    for (long k = 0; k < y; k++)
                        frag[(size_t)(x + k)] = fu.strand_b > 0 ? gb[fu.pos_b + k] : comp(gb[fu.pos_b - k]);
                    kind = 2;
                }
            }
            if (kind == 1) {
                for (long k = 0; k < f; k += 32) {
                    uint64_t xw = r.next();
                    for (int j = 0; j < 32 && k + j < f; j++) frag[(size_t)(k + j)] = BASES[(xw >> (2 * j)) & 3];
                }
            }
            /* sequencing strand */
            bool flip = (r.next() & 1) != 0;
            uint8_t* s1 = seq1 + p * (uint64_t)L;
            uint8_t* s2 = seq2 + p * (uint64_t)L;
            uint8_t* q1 = qual1 + p * (uint64_t)L;
            uint8_t* q2 = qual2 + p * (uint64_t)L;
            for (int k = 0; k < L; k++) {
                uint8_t a = frag[(size_t)k], b = comp(frag[(size_t)(f - 1 - k)]);
                if (flip) { uint8_t t = a; a = b; b = t; }
                s1[k] = a;
                s2[k] = b;
            }
            /* errors + qualities */
            for (int m = 0; m < 2; m++) {
                uint8_t* s = m ? s2 : s1;
                uint8_t* q = m ? q2 : q1;
                for (int k = 0; k < L; k++) {
                    double e = r.uni();
                    if (e < cfg->n_rate) s[k] = 'N';
                    else if (e < cfg->n_rate + cfg->sub_rate) {
                        const char* at = (const char*)memchr("ACGT", s[k], 4);
                        uint64_t idx = at ? (uint64_t)(at - "ACGT") : 0;
                        uint8_t nb = BASES[(idx + 1 + r.below(3)) & 3]; /* always a different base */
                        s[k] = nb;
                    }
                    double qv = r.uni();
                    q[k] = qv < 0.90 ? 'E' : (qv < 0.97 ? 'A' : '/');
                }
            }
            if (kind_out) kind_out[p] = kind;
        }
    };
    if (threads <= 1 || count < 4096) {
        work(0, count);
    } else {
        std::vector<std::thread> th;
        uint64_t per = (count + (uint64_t)threads - 1) / (uint64_t)threads;
        for (int t = 0; t < threads; t++) {
            uint64_t lo = (uint64_t)t * per, hi = lo + per > count ? count : lo + per;
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& x : th) x.join();
    }
}

} /* extern "C" */
