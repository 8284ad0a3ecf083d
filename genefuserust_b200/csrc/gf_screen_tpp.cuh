/*
 * gf_screen_tpp.cuh — thread-per-pair building blocks of the screen (32 pairs per warp, reads up to 256 bases), used by
 * k_prep (gf_screen_split.cuh) and by the fast_merge parity hook:
 *
 *  - each thread keeps its planes in a private shared-memory column: word k of thread t lives at [k][t], so any
 *    data-dependent index still hits the thread's own bank (never a conflict);
 *  - reads are converted 32 bases per step: one aligned 256-bit sector load, two 16-base SWAR blocks (block16 below);
 *  - reverse_complement(R2) is produced directly by walking R2 backwards;
 *  - qualities are NOT converted: fast_merge (read.rs:313-440) only looks at them where R1 and rc(R2) disagree inside a
 *    candidate overlap that has <= 2 mismatches, so the two bytes are fetched from global memory only then.
 *
 * Included by gf_map.cu (uses lowmask / zero_bytes defined there).
 */
#pragma once

namespace tpp {

using swar::block16;
using swar::fill16;

/* private column layout (word offsets) for reads of up to 32*W bases: W = 5 (<= 160 bases, 48 words per thread)
 * or W = 8 (<= 256 bases, 72 words per thread): R1 forward planes and the planes of rc(R2) */
template <int W>
struct Lay {
    static constexpr int R1LO = 0, R1HI = R1LO + W + 1, R1V = R1HI + W + 1, R1N = R1V + W + 1;
    static constexpr int C2LO = R1N + W + 1, C2HI = C2LO + W + 1, C2V = C2HI + W + 1, VCS = C2V + W + 1;
    static constexpr int NWORDS = VCS + W + 1;
};
constexpr int WARPS = 4;

struct Col { /* a thread's private column */
    uint32_t* base; /* &smem[warp][0][lane] */
    __device__ __forceinline__ uint32_t& operator()(int arr, int k) const { return base[(arr + k) * 32]; }
    __device__ __forceinline__ uint32_t fs(int arr, uint32_t bitpos) const {
        uint32_t w = bitpos >> 5;
        return __funnelshift_r(base[(arr + (int)w) * 32], base[(arr + (int)w + 1) * 32], bitpos & 31u);
    }
    __device__ __forceinline__ uint32_t win(int arr, int pos) const {
        if (pos >= 0) return fs(arr, (uint32_t)pos);
        if (pos > -32) return base[arr * 32] << (-pos);
        return 0u;
    }
};

/* aligned 32-bit load; bytes outside [lo, hi) read as 0 */
__device__ __forceinline__ uint32_t load4_guarded(const uint8_t* p, const uint8_t* lo, const uint8_t* hi) {
    if (p >= lo && p + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(p));
    uint32_t w = 0;
    for (int t = 0; t < 4; t++)
        if (p + t >= lo && p + t < hi) w |= (uint32_t)__ldg(p + t) << (8 * t);
    return w;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* the block straddles an end of the arena (first / last read of a batch only): byte by byte, kept out of line */
__device__ __noinline__ uint4 load16_slow(const uint8_t* p, const uint8_t* lo, const uint8_t* hi) {
    return make_uint4(load4_guarded(p, lo, hi), load4_guarded(p + 4, lo, hi), load4_guarded(p + 8, lo, hi),
                      load4_guarded(p + 12, lo, hi));
}
/* 32-aligned 256-bit load (one full sector per request), streaming; `span_inside` = the caller has checked that the whole
 * aligned span of the read lies inside the arena (false only for the first / last reads of a batch) */
__device__ __forceinline__ void load32_span(const uint8_t* p, bool span_inside, const uint8_t* lo, const uint8_t* hi,
                                            unsigned long long pol_stream, uint4* a, uint4* b) {
    if (span_inside) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w)
                     : "l"(p), "l"(pol_stream));
        return;
    }
    *a = load16_slow(p, lo, hi);
    *b = load16_slow(p + 16, lo, hi);
}
__device__ __forceinline__ unsigned long long make_policy_normal() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long make_policy_stream() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint32_t tailmask(int rem) { /* bits of a plane word that lie inside the read */
    return rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? (1u << rem) - 1u : 0u);
}

/* R1 -> forward planes lo, hi, valid (upper-case ACGT), N */
template <int W>
__device__ __forceinline__ void convert_r1(const Col& c, const uint8_t* seq, int len, const uint8_t* lo, const uint8_t* hi,
                                           unsigned long long pol_stream) {
    const uint32_t a = (uint32_t)((uintptr_t)seq & 31u);
    const uint8_t* bp = seq - a; /* 32-aligned, advanced by 32 per step */
    const int nu = len + (int)a;                      /* u positions in use: [a, nu) */
    const int nwords = len > 0 ? (nu + 31) >> 5 : 0;  /* U words, <= W + 1 */
    uint32_t plo = 0, phi = 0, pv = 0, pn = 0;        /* previous U word */
    const bool inside = bp >= lo && bp + 32 * nwords <= hi;
    uint4 A = fill16(), B = fill16();
    if (nwords > 0) load32_span(bp, inside, lo, hi, pol_stream, &A, &B);
#pragma unroll 1
    for (int m = 0; m < nwords; m++) {
        uint4 nA = fill16(), nB = fill16();
        bp += 32;
        if (m + 1 < nwords) load32_span(bp, inside, lo, hi, pol_stream, &nA, &nB);
        uint32_t l0, h0, v0, n0, l1, h1, v1, n1;
        block16<false>(A, &l0, &h0, &v0, &n0);
        block16<false>(B, &l1, &h1, &v1, &n1);
        const uint32_t ulo = __byte_perm(l0, l1, 0x5410u), uhi = __byte_perm(h0, h1, 0x5410u);
        const uint32_t uv = __byte_perm(v0, v1, 0x5410u), un = __byte_perm(n0, n1, 0x5410u);
        if (m > 0) {
            const uint32_t tm = tailmask(len - 32 * (m - 1));
            c(Lay<W>::R1LO, m - 1) = __funnelshift_r(plo, ulo, a) & tm; c(Lay<W>::R1HI, m - 1) = __funnelshift_r(phi, uhi, a) & tm;
            c(Lay<W>::R1V, m - 1) = __funnelshift_r(pv, uv, a) & tm;    c(Lay<W>::R1N, m - 1) = __funnelshift_r(pn, un, a) & tm;
        }
        plo = ulo; phi = uhi; pv = uv; pn = un;
        A = nA; B = nB;
    }
    if (nwords > 0) {
        const uint32_t tm = tailmask(len - 32 * (nwords - 1));
        c(Lay<W>::R1LO, nwords - 1) = (plo >> a) & tm; c(Lay<W>::R1HI, nwords - 1) = (phi >> a) & tm;
        c(Lay<W>::R1V, nwords - 1) = (pv >> a) & tm;   c(Lay<W>::R1N, nwords - 1) = (pn >> a) & tm;
    }
    for (int w = nwords; w <= W; w++) { c(Lay<W>::R1LO, w) = 0; c(Lay<W>::R1HI, w) = 0; c(Lay<W>::R1V, w) = 0; c(Lay<W>::R1N, w) = 0; }
}

/* R2 -> planes of reverse_complement(R2) (case-insensitive, sequence.rs:52-60) + the case-sensitive validity of
 * the same bases (VCS, in rc orientation), produced by walking R2 from its last byte downwards: the 32 bytes below
 * `top - 32 m` give rc positions u' = 32 m .. 32 m + 31 in reversed bit order, so one BREV per plane word */
template <int W>
__device__ __forceinline__ void convert_r2_rc(const Col& c, const uint8_t* seq, int len, const uint8_t* lo, const uint8_t* hi,
                                              unsigned long long pol_stream) {
    const uint8_t* end = seq + len;
    const uint32_t pad = (uint32_t)((32u - ((uintptr_t)end & 31u)) & 31u); /* garbage bytes above the last base */
    const uint8_t* top = end + pad;           /* 32-aligned, lowered by 32 per step */
    const int nu = len + (int)pad;
    const int nwords = len > 0 ? (nu + 31) >> 5 : 0;
    uint32_t plo = 0, phi = 0, pv = 0, pc = 0;
    const bool inside = top <= hi && top - 32 * nwords >= lo;
    uint4 A = fill16(), B = fill16();         /* A: rc positions u' 0..15 of the word (upper 16 bytes), B: 16..31 */
    if (nwords > 0) load32_span(top - 32, inside, lo, hi, pol_stream, &B, &A);
#pragma unroll 1
    for (int m = 0; m < nwords; m++) {
        uint4 nA = fill16(), nB = fill16();
        top -= 32;
        if (m + 1 < nwords) load32_span(top - 32, inside, lo, hi, pol_stream, &nB, &nA);
        uint32_t l0, h0, v0, e0, l1, h1, v1, e1;
        block16<true>(A, &l0, &h0, &v0, &e0);
        block16<true>(B, &l1, &h1, &v1, &e1);
        const uint32_t ulo = __brev(__byte_perm(l1, l0, 0x5410u)), uhi = __brev(__byte_perm(h1, h0, 0x5410u));
        const uint32_t uv = __brev(__byte_perm(v1, v0, 0x5410u)), uc = __brev(__byte_perm(e1, e0, 0x5410u));
        if (m > 0) {
            const uint32_t tm = tailmask(len - 32 * (m - 1));
            const uint32_t sv = __funnelshift_r(pv, uv, pad) & tm;
            c(Lay<W>::C2LO, m - 1) = ~__funnelshift_r(plo, ulo, pad) & sv; /* complement = code ^ 1 */
            c(Lay<W>::C2HI, m - 1) = __funnelshift_r(phi, uhi, pad) & sv;
            c(Lay<W>::C2V, m - 1) = sv;
            c(Lay<W>::VCS, m - 1) = __funnelshift_r(pc, uc, pad) & tm;
        }
        plo = ulo; phi = uhi; pv = uv; pc = uc;
        A = nA; B = nB;
    }
    if (nwords > 0) {
        const uint32_t tm = tailmask(len - 32 * (nwords - 1));
        const uint32_t sv = (pv >> pad) & tm;
        c(Lay<W>::C2LO, nwords - 1) = ~(plo >> pad) & sv; c(Lay<W>::C2HI, nwords - 1) = (phi >> pad) & sv;
        c(Lay<W>::C2V, nwords - 1) = sv;                  c(Lay<W>::VCS, nwords - 1) = (pc >> pad) & tm;
    }
    for (int w = nwords; w <= W; w++) { c(Lay<W>::C2LO, w) = 0; c(Lay<W>::C2HI, w) = 0; c(Lay<W>::C2V, w) = 0; c(Lay<W>::VCS, w) = 0; }
}

/* ---- packed upload: the host built the forward planes (gf_pack.cpp), nothing is converted here ---- */
/* R1: nw words of lo, nw of hi; exception words (valid = upper-case ACGT, aux = 'N') only when the read has any */
template <int W>
__device__ __forceinline__ void load_r1_packed(const Col& c, const uint32_t* __restrict__ w, const uint32_t* __restrict__ x, int len) {
    const int nw = (len + 31) >> 5;
    for (int k = 0; k < nw; k++) {
        c(Lay<W>::R1LO, k) = __ldg(w + k);
        c(Lay<W>::R1HI, k) = __ldg(w + nw + k);
        c(Lay<W>::R1V, k) = x ? __ldg(x + k) : tailmask(len - 32 * k);
        c(Lay<W>::R1N, k) = x ? __ldg(x + nw + k) : 0u;
    }
    for (int k = nw; k <= W; k++) { c(Lay<W>::R1LO, k) = 0; c(Lay<W>::R1HI, k) = 0; c(Lay<W>::R1V, k) = 0; c(Lay<W>::R1N, k) = 0; }
}
/* 32 bits of a forward plane (nw words at p) starting at bit `pos` (may be negative); bits outside the plane read as 0 */
__device__ __forceinline__ uint32_t plane_win(const uint32_t* __restrict__ p, int nw, int pos) {
    if (pos <= -32) return 0u;
    if (pos < 0) return __ldg(p) << (-pos);
    const int wi = pos >> 5;
    const uint32_t a = __ldg(p + wi), b = wi + 1 < nw ? __ldg(p + wi + 1) : 0u;
    return __funnelshift_r(a, b, (uint32_t)pos & 31u);
}
/* R2: forward planes -> the planes of reverse_complement(R2): rc position u' = len - 1 - u, so word k of an rc plane is the
 * bit-reversed forward window at len - 32 (k + 1); complement = code ^ 1 = the low plane inverted.  Exception words: valid
 * = ACGT in either case, aux = upper-case ACGT (VCS) */
template <int W>
__device__ __forceinline__ void load_r2_rc_packed(const Col& c, const uint32_t* __restrict__ w, const uint32_t* __restrict__ x, int len) {
    const int nw = (len + 31) >> 5;
    for (int k = 0; k < nw; k++) {
        const int pos = len - 32 * (k + 1);
        const uint32_t tm = tailmask(len - 32 * k);
        const uint32_t sv = x ? __brev(plane_win(x, nw, pos)) : tm;
        c(Lay<W>::C2LO, k) = ~__brev(plane_win(w, nw, pos)) & sv;
        c(Lay<W>::C2HI, k) = __brev(plane_win(w + nw, nw, pos)) & sv;
        c(Lay<W>::C2V, k) = sv;
        c(Lay<W>::VCS, k) = x ? __brev(plane_win(x + nw, nw, pos)) : tm;
    }
    for (int k = nw; k <= W; k++) { c(Lay<W>::C2LO, k) = 0; c(Lay<W>::C2HI, k) = 0; c(Lay<W>::C2V, k) = 0; c(Lay<W>::VCS, k) = 0; }
}

/* mismatch mask of overlap chunk k for overlap length olen (read.rs:346) */
template <int W>
__device__ __forceinline__ uint32_t overlap_mism(const Col& c, int offset, int olen, int len2, int k) {
    uint32_t bp = (uint32_t)(offset + 32 * k);
    uint32_t alo = c.fs(Lay<W>::R1LO, bp), ahi = c.fs(Lay<W>::R1HI, bp), av = c.fs(Lay<W>::R1V, bp), an = c.fs(Lay<W>::R1N, bp);
    uint32_t cv = c(Lay<W>::C2V, k);
    uint32_t cn = lowmask(len2 - 32 * k) & ~cv; /* rc(R2) holds 'N' wherever R2 is not ACGT/acgt */
    return ((alo ^ c(Lay<W>::C2LO, k)) | (ahi ^ c(Lay<W>::C2HI, k)) | (an ^ cn) | (~av & ~an)) & lowmask(olen - 32 * k);
}
/* "one is >= Q30 and the other <= Q15" (read.rs:349-354) for overlap position i */
__device__ __forceinline__ bool low_qual_pair(const uint8_t* q1, const uint8_t* q2, int offset, int len2, int i, bool* r1_wins) {
    uint32_t a = __ldg(q1 + offset + i), b = __ldg(q2 + (len2 - 1 - i));
    *r1_wins = a >= '?' && b <= '0';
    return *r1_wins || (a <= '0' && b >= '?');
}
/* the full test of one overlap length (read.rs:339-367): <= 2 mismatches, each one a "low quality" pair.
 * *r1_bits: bit i set <=> at the i-th mismatch (in position order) the R1 base wins (q1 >= '?' and q2 <= '0',
 * read.rs:401-428) — kept so that the merged read can be built without fetching the quality bytes a second time
 * (they may live in pinned host memory) */
template <int W>
__device__ __forceinline__ bool overlap_passes(const Col& c, int len1, int len2, int o, const uint8_t* q1, const uint8_t* q2, int* diff_out,
                                               uint32_t* r1_bits) {
    const int offset = len1 - o;
    int cnt = 0;
    uint32_t bits = 0, nm = 0;
    for (int k = 0; 32 * k < o; k++) {
        uint32_t mism = overlap_mism<W>(c, offset, o, len2, k);
        if (!mism) continue;
        cnt += __popc(mism);
        if (cnt > 2) return false;
        while (mism) {
            int b = __ffs(mism) - 1;
            mism &= mism - 1;
            bool r1w;
            if (!low_qual_pair(q1, q2, offset, len2, 32 * k + b, &r1w)) return false;
            bits |= (r1w ? 1u : 0u) << nm;
            nm++;
        }
    }
    *diff_out = cnt;
    *r1_bits = bits;
    return true;
}
/* smallest passing overlap length (read.rs:323-367) or -1.
 * Cheap reject per overlap length o >= 32: the low code-bit plane of R1 at offset len1 - o against the first 32 bases of rc(R2)
 * must differ in <= 2 positions (o = 30, 31: all 30 / 31 positions).  The R1 window slides by one bit per o, so the plane words
 * are loaded once per 32 overlap lengths and every o costs funnel shift + xor + popc + 2 (candidate bit shifted into a mask, no
 * branch).
 * The scan runs over ALL overlap lengths first and only notes the candidates (the first two, in increasing o); the full test
 * (overlap_passes: plane windows, quality bytes from global / pinned host memory) comes afterwards, so that the lanes of a warp
 * that have a candidate — every pair that merges has one, at its own o — run it side by side instead of one after the other
 * whenever their candidate happens to show up in the scan.  A third candidate (low-complexity reads) is left to
 * find_overlap_from, the one-by-one form of the same search. */
template <int W>
__device__ __noinline__ int find_overlap_from(const Col& c, int len1, int len2, int o_min, const uint8_t* q1, const uint8_t* q2, int* diff_out,
                                              uint32_t* r1_bits) {
    const int minlen = min(len1, len2);
    const uint32_t c0 = c(Lay<W>::C2LO, 0);
    for (int o = max(32, o_min); o <= minlen; o++) {
        const uint32_t x = c.fs(Lay<W>::R1LO, (uint32_t)(len1 - o)) ^ c0;
        if (__popc(x) <= 2 && overlap_passes<W>(c, len1, len2, o, q1, q2, diff_out, r1_bits)) return o;
    }
    *diff_out = 0;
    return -1;
}
template <int W>
__device__ __forceinline__ int find_overlap(const Col& c, int len1, int len2, const uint8_t* q1, const uint8_t* q2, int* diff_out,
                                            uint32_t* r1_bits) {
    const int minlen = min(len1, len2);
    *diff_out = 0;
    *r1_bits = 0;
    int o1 = 0, o2 = 0; /* the first two candidates */
    bool more = false;
    const uint32_t c0 = c(Lay<W>::C2LO, 0);
    /* o = 30, 31: fewer than 32 positions */
    for (int o = 30; o <= min(31, minlen); o++) {
        const uint32_t x = (c.fs(Lay<W>::R1LO, (uint32_t)(len1 - o)) ^ c0) & lowmask(o);
        if (__popc(x) <= 2) { if (!o1) o1 = o; else o2 = o; }
    }
    if (minlen >= 32) {
        const int w_hi = (len1 - 32) >> 5, w_lo = (len1 - minlen) >> 5;
        uint32_t b = c(Lay<W>::R1LO, w_hi + 1);
#pragma unroll 1
        for (int w = w_hi; w >= w_lo; w--) {
            const uint32_t a = c(Lay<W>::R1LO, w);
            uint32_t cand = 0; /* bit 31 - j <-> offset 32 w + 31 - j, i.e. o = o_base + j */
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t x = __funnelshift_r(a, b, 31 - j) ^ c0;
                cand = __funnelshift_l((uint32_t)(__popc(x) - 3), cand, 1);
            }
            b = a;
            const int o_base = len1 - 32 * w - 31;
            /* keep 32 <= o <= minlen */
            const int j_lo = max(0, 32 - o_base), j_hi = min(31, minlen - o_base);
            if (j_hi < j_lo) continue;
            cand &= (0xFFFFFFFFu >> j_lo) & (0xFFFFFFFFu << (31 - j_hi));
            while (cand) { /* rare: candidates in increasing o */
                const int j = __clz(cand);
                cand &= ~(0x80000000u >> j);
                if (!o1) o1 = o_base + j;
                else if (!o2) o2 = o_base + j;
                else { more = true; cand = 0; }
            }
        }
    }
    if (o1 && overlap_passes<W>(c, len1, len2, o1, q1, q2, diff_out, r1_bits)) return o1;
    if (o2 && overlap_passes<W>(c, len1, len2, o2, q1, q2, diff_out, r1_bits)) return o2;
    if (more) return find_overlap_from<W>(c, len1, len2, o2 + 1, q1, q2, diff_out, r1_bits);
    *diff_out = 0;
    return -1;
}
/* parity hook: fast_merge only, thread per pair */
template <int W>
__global__ void __launch_bounds__(WARPS * 32) k_merge_only_tpp(GfDevBatch B, gf_merge_info* __restrict__ out,
                                                                GfMapCounters* counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t lane = gf_lane(), wib = threadIdx.x >> 5;
    Col c;
    c.base = sm + (size_t)wib * Lay<W>::NWORDS * 32 + lane;
    const uint64_t stride = (uint64_t)gridDim.x * WARPS * 32;
    for (uint64_t p = ((uint64_t)blockIdx.x * WARPS + wib) * 32 + lane; p < B.n; p += stride) {
        const uint64_t o1 = B.s1[p], o2 = B.s2[p];
        const int len1 = (int)(B.e1[p] - o1), len2 = (int)(B.e2[p] - o2);
        if (len1 > 32 * W || len2 > 32 * W) { atomicOr(&counters->error_flags, 1u); continue; }
        convert_r1<W>(c, B.seq1 + (o1 - B.base1), len1, B.seq1, B.seq1 + B.bytes1, make_policy_stream());
        convert_r2_rc<W>(c, B.seq2 + (o2 - B.base2), len2, B.seq2, B.seq2 + B.bytes2, make_policy_stream());
        int diff = 0;
        uint32_t r1_bits_unused;
        int olen = find_overlap<W>(c, len1, len2, B.qual1 + (B.qs1[p] - B.base1), B.qual2 + (B.qs2[p] - B.base2), &diff, &r1_bits_unused);
        gf_merge_info mi;
        mi.merged = olen >= 0;
        mi.olen = olen >= 0 ? olen : 0;
        mi.diff = diff;
        mi.merged_len = olen >= 0 ? len1 - olen + len2 : 0;
        out[p] = mi;
    }
}

}  // namespace tpp
