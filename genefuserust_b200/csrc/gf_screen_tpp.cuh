/*
 * gf_screen_tpp.cuh — screen kernel v3: THREAD per pair (32 pairs per warp), reads up to 256 bases.
 *
 * Same decisions as k_screen v2 (fast_merge, read.rs:313-440; conservative first pass of Indexer::map_read,
 * indexer.rs:252-360 — see the bound in gf_map.cu), but every step that was "one warp, 5-9 useful lanes" in the
 * warp-per-pair kernel is now 32 pairs in lock step, which cuts the warp instructions per pair ~4x
 * (profiles/r01_screen_v2c_by_function.txt: the v2 kernel is issue bound at ~2070 instructions per pair).
 *
 *  - each thread keeps its planes in a private shared-memory column: word k of thread t lives at [k][t], so any
 *    data-dependent index still hits the thread's own bank (never a conflict);
 *  - reads are converted 32 bases per step: one aligned 256-bit sector load, two 16-base SWAR blocks (block16 below);
 *  - reverse_complement(R2) is produced directly by walking R2 backwards;
 *  - qualities are NOT converted: fast_merge only looks at them where R1 and rc(R2) disagree inside a candidate
 *    overlap that has <= 2 mismatches, so the two bytes are fetched from global memory only then;
 *  - the gene planes along the seed diagonal are streamed through registers (no staging).
 *
 * Included by gf_map.cu (uses ScreenParams, gather4/zero_bytes/classify4/run16/ldg_* defined there).
 */
#pragma once

namespace tpp {

/* private column layout (word offsets) for reads of up to 32*W bases: W = 5 (<= 160 bases, 81 words per thread)
 * or W = 8 (<= 256 bases, 123 words per thread) */
template <int W>
struct Lay {
    static constexpr int WM = 2 * W; /* merged read */
    static constexpr int R1LO = 0, R1HI = R1LO + W + 1, R1V = R1HI + W + 1, R1N = R1V + W + 1;
    static constexpr int C2LO = R1N + W + 1, C2HI = C2LO + W + 1, C2V = C2HI + W + 1, VCS = C2V + W + 1;
    static constexpr int MLO = VCS + W + 1, MHI = MLO + WM + 1, MV = MHI + WM + 1;
    static constexpr int NWORDS = MV + WM + 1;
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

/* Both converters read one aligned 32-byte sector (two 16-byte blocks) per step and turn each block into 16 plane bits with a handful of SWAR
 * operations (no per-base work):
 *   - the code bits (bit 2 / bit 1 of the ASCII byte = A0 T1 C2 G3) of 8 bytes are gathered by ONE multiply:
 *     z = low nibbles of word 0 | low nibbles of word 1 << 4, (z & 0x44444444) * 0x00408102 has the eight bit-2 values in
 *     base order in its top byte (0x00810204 for bit 1); the partial products never collide, so there are no carries;
 *   - validity: the expected letter is looked up with PRMT from the low 3 bits of each byte (A 001, C 011, T 100, G 111)
 *     and xor-ed with the byte; a block whose 16 differences are all zero takes the fast path (valid = 0xFFFF);
 *   - misalignment a (0..31) of the read w.r.t. the sectors is removed in the bit domain: planes are built at bit position
 *     u = p + a and aligned word w = funnelshift(U[w], U[w+1], a). */
__device__ __forceinline__ uint32_t expect4(uint32_t x) {
    const uint32_t t = x & 0x07070707u;
    const uint32_t u = t | (t >> 4);
    return __byte_perm(0x43414141u, 0x47414154u, __byte_perm(u, 0u, 0x4420u));
}
/* One 16-byte block -> 16 plane bits in the LOW half of each result (the upper halves are garbage):
 *   lo / hi = code bits, v = valid (ACGT; either case when CI), ex = !CI: the byte is 'N';  CI: valid AND upper case */
template <bool CI>
__device__ __forceinline__ void block16(const uint4& x, uint32_t* lo, uint32_t* hi, uint32_t* v, uint32_t* ex) {
    const uint32_t z0 = (x.x & 0x0F0F0F0Fu) | ((x.y << 4) & 0xF0F0F0F0u);
    const uint32_t z1 = (x.z & 0x0F0F0F0Fu) | ((x.w << 4) & 0xF0F0F0F0u);
    uint32_t l = __byte_perm((z0 & 0x44444444u) * 0x00408102u, (z1 & 0x44444444u) * 0x00408102u, 0x7373u);
    uint32_t h = __byte_perm((z0 & 0x22222222u) * 0x00810204u, (z1 & 0x22222222u) * 0x00810204u, 0x7373u);
    constexpr uint32_t CM = CI ? 0xDFDFDFDFu : 0xFFFFFFFFu;
    const uint32_t d0 = (x.x ^ expect4(x.x)) & CM, d1 = (x.y ^ expect4(x.y)) & CM;
    const uint32_t d2 = (x.z ^ expect4(x.z)) & CM, d3 = (x.w ^ expect4(x.w)) & CM;
    uint32_t bad = d0 | d1 | d2 | d3;
    if (CI) bad |= (x.x | x.y | x.z | x.w) & 0x20202020u; /* a lower-case letter: ex differs from v */
    uint32_t vv = 0xFFFFu, e = CI ? 0xFFFFu : 0u;
    if (bad) { /* rare: N, lower case, bytes outside the arena */
        const uint32_t y0 = (zero_bytes(d0) >> 5) | (zero_bytes(d1) >> 1);
        const uint32_t y1 = (zero_bytes(d2) >> 5) | (zero_bytes(d3) >> 1);
        vv = __byte_perm(y0 * 0x00408102u, y1 * 0x00408102u, 0x7373u);
        if (CI) {
            const uint32_t w0 = ((x.x >> 3) & 0x04040404u) | ((x.y << 1) & 0x40404040u);
            const uint32_t w1 = ((x.z >> 3) & 0x04040404u) | ((x.w << 1) & 0x40404040u);
            e = vv & ~__byte_perm(w0 * 0x00408102u, w1 * 0x00408102u, 0x7373u);
        } else {
            const uint32_t n0 = (zero_bytes(x.x ^ 0x4E4E4E4Eu) >> 5) | (zero_bytes(x.y ^ 0x4E4E4E4Eu) >> 1);
            const uint32_t n1 = (zero_bytes(x.z ^ 0x4E4E4E4Eu) >> 5) | (zero_bytes(x.w ^ 0x4E4E4E4Eu) >> 1);
            e = __byte_perm(n0 * 0x00408102u, n1 * 0x00408102u, 0x7373u);
        }
        l &= vv; h &= vv;
    }
    *lo = l; *hi = h; *v = vv; *ex = e;
}
__device__ __forceinline__ uint4 fill16() { return make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u); }
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
 * Cheap reject per overlap length o: the low code-bit plane of R1 at offset len1 - o against the first 32 bases of rc(R2)
 * must differ in <= 2 positions.  The R1 window slides by one bit per o, so the plane words are loaded once per 32 overlap
 * lengths and every o costs funnel shift + xor + popc + 2 (candidate bit shifted into a mask, no branch). */
template <int W>
__device__ __forceinline__ int find_overlap(const Col& c, int len1, int len2, const uint8_t* q1, const uint8_t* q2, int* diff_out,
                                            uint32_t* r1_bits) {
    const int minlen = min(len1, len2);
    *diff_out = 0;
    *r1_bits = 0;
    /* o = 30, 31: fewer than 32 positions */
    for (int o = 30; o <= min(31, minlen); o++)
        if (overlap_passes<W>(c, len1, len2, o, q1, q2, diff_out, r1_bits)) return o;
    if (minlen < 32) return -1;
    const uint32_t c0 = c(Lay<W>::C2LO, 0);
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
            if (overlap_passes<W>(c, len1, len2, o_base + j, q1, q2, diff_out, r1_bits)) return o_base + j;
        }
    }
    *diff_out = 0;
    return -1;
}
/* merged read planes into M (read.rs:369-428) */
template <int W>
__device__ __forceinline__ int build_merged(const Col& c, int len1, int len2, int olen, const uint8_t* q1, const uint8_t* q2) {
    const int offset = len1 - olen, mlen = offset + len2, nwm = (mlen + 31) >> 5;
    for (int w = 0; w <= Lay<W>::WM; w++) {
        uint32_t lo = 0, hi = 0, v = 0;
        if (w < nwm) {
            int pos0 = 32 * w;
            if (pos0 < offset) {
                uint32_t m = lowmask(offset - pos0);
                lo = c(Lay<W>::R1LO, w) & m; hi = c(Lay<W>::R1HI, w) & m; v = c(Lay<W>::R1V, w) & m;
            }
            int j0 = pos0 - offset;
            if (j0 > -32) { lo |= c.win(Lay<W>::C2LO, j0); hi |= c.win(Lay<W>::C2HI, j0); v |= c.win(Lay<W>::C2V, j0); }
        }
        c(Lay<W>::MLO, w) = lo; c(Lay<W>::MHI, w) = hi; c(Lay<W>::MV, w) = v;
    }
    /* overlap mismatches keep the R1 base iff q1 >= '?' and q2 <= '0' (at most 2 positions) */
    for (int k = 0; 32 * k < olen; k++) {
        uint32_t mism = overlap_mism<W>(c, offset, olen, len2, k);
        while (mism) {
            int b = __ffs(mism) - 1;
            mism &= mism - 1;
            bool r1w;
            low_qual_pair(q1, q2, offset, len2, 32 * k + b, &r1w);
            if (r1w) {
                int pos = offset + 32 * k + b;
                uint32_t bit = 1u << (pos & 31), w = (uint32_t)pos >> 5;
                c(Lay<W>::MLO, w) = (c(Lay<W>::MLO, w) & ~bit) | (c(Lay<W>::R1LO, w) & bit);
                c(Lay<W>::MHI, w) = (c(Lay<W>::MHI, w) & ~bit) | (c(Lay<W>::R1HI, w) & bit);
                c(Lay<W>::MV, w) = (c(Lay<W>::MV, w) & ~bit) | (c(Lay<W>::R1V, w) & bit);
            }
        }
    }
    return mlen;
}
/* forward planes of R2 (upper-case validity) into M, from the rc planes */
template <int W>
__device__ __forceinline__ void build_r2_forward(const Col& c, int len2) {
    const int nw = (len2 + 31) >> 5;
    for (int w = 0; w <= Lay<W>::WM; w++) {
        uint32_t lo = 0, hi = 0, v = 0;
        if (w < nw) {
            int pos = len2 - 32 * w - 32;
            v = __brev(c.win(Lay<W>::VCS, pos));
            lo = ~__brev(c.win(Lay<W>::C2LO, pos)) & v;
            hi = __brev(c.win(Lay<W>::C2HI, pos)) & v;
        }
        c(Lay<W>::MLO, w) = lo; c(Lay<W>::MHI, w) = hi; c(Lay<W>::MV, w) = v;
    }
}

__device__ __forceinline__ uint32_t filter_sites(const GfDevIndex& ix, uint32_t key, unsigned long long pol, uint32_t max_sites) {
    return gf_filter_sites(ix, ldg_filter(ix.filter + gf_filter_word(key, ix.filter_words), pol), key, max_sites);
}

/* conservative first pass over one sequence whose planes start at (LO, HI, V) in the private column */
template <class COL>
__device__ __forceinline__ bool screen_sequence(const GfDevIndex& ix, const COL& c, int LO, int HI, int V, int len,
                                                int need_total, int need_minor, unsigned long long pol) {
    const int nprobe = len >= 16 ? ((len - 16) >> 1) + 1 : 0;
    if (need_total <= 0 || need_minor <= 0) return nprobe > 0 || need_total <= 0;
    if (nprobe == 0) return false;
    const int nch = (len + 31) >> 5;

    /* 1. seed: first of 8 spread k-mers that the filter calls unique and the HBM table confirms */
    uint32_t seed_val = GF_EMPTY_VAL;
    int seed_i = 0;
#pragma unroll 1
    for (int s = 0; s < 8; s++) {
        int i = (int)(((long long)s * nprobe) >> 3) * 2;
        if ((c.fs(V, (uint32_t)i) & 0xFFFFu) != 0xFFFFu) continue;
        uint32_t key = ((c.fs(HI, (uint32_t)i) & 0xFFFFu) << 16) | (c.fs(LO, (uint32_t)i) & 0xFFFFu);
        if (filter_sites(ix, key, pol, 2u) != 1u) continue;
        uint32_t val = gf_table_find(ix, key);
        if (val != GF_EMPTY_VAL && (val >> 30) == GF_KIND_UNIQUE) { seed_val = val; seed_i = i; break; }
    }

    /* 2. walk the read in 32-base chunks.  With a seed: compare with the gene planes along the diagonal (votes of
     * matching indexed windows are exact).  Every valid even offset the diagonal does not explain — all of them
     * when there is no seed — gets a filter probe (upper bound of its votes), four probes in flight at a time. */
    const bool seeded = seed_val != GF_EMPTY_VAL;
    const bool rc = seeded && (seed_val & GF_SITE_STRAND) != 0;
    const uint32_t goff = seed_val & GF_SITE_GOFF_MASK;
    const uint32_t D = rc ? goff + (uint32_t)seed_i - (uint32_t)len + 1u : goff - (uint32_t)seed_i;
    const uint32_t parity = (rc && (len & 1)) ? 0xAAAAAAAAu : 0x55555555u;
    const uint32_t wbase = seeded ? (D >> 5) : 0u, sh = D & 31u;
    const uint32_t* gc = rc ? ix.g_cr : ix.g_cf;
    /* read chunk k in the orientation of the comparison (reverse complement for a reverse-strand seed) */
    auto read_chunk = [&](int k, uint32_t* lo, uint32_t* hi, uint32_t* v) {
        if (k >= nch) { *lo = *hi = *v = 0; return; }
        if (!rc) { *lo = c(LO, k); *hi = c(HI, k); *v = c(V, k); return; }
        int pos = len - 32 * k - 32;
        uint32_t vv = __brev(c.win(V, pos));
        *v = vv;
        *lo = ~__brev(c.win(LO, pos)) & vv;
        *hi = __brev(c.win(HI, pos));
    };
    /* gene words, sliding: g*0 = word (wbase+k+1) after chunk k has been formed */
    uint32_t glo0 = 0, ghi0 = 0, gv0 = 0, gca0 = 0, gcb0 = 0, gcc0 = 0;
    uint32_t e_cur = 0, cnt_a = 0, cnt_b = 0, cnt_c = 0; /* equality word and site-count bits of chunk k */
    uint32_t lo_cur, hi_cur, v_cur;
    read_chunk(0, &lo_cur, &hi_cur, &v_cur);
    if (seeded) {
        glo0 = ldg_plane(ix.g_lo + wbase, pol); ghi0 = ldg_plane(ix.g_hi + wbase, pol); gv0 = ldg_plane(ix.g_v + wbase, pol);
        gca0 = ldg_plane(gc + wbase, pol); gcb0 = ldg_plane(gc + ix.g_cstride + wbase, pol);
        gcc0 = ldg_plane(gc + 2 * ix.g_cstride + wbase, pol);
        uint32_t glo1 = ldg_plane(ix.g_lo + wbase + 1, pol), ghi1 = ldg_plane(ix.g_hi + wbase + 1, pol),
                 gv1 = ldg_plane(ix.g_v + wbase + 1, pol);
        uint32_t gca1 = ldg_plane(gc + wbase + 1, pol), gcb1 = ldg_plane(gc + ix.g_cstride + wbase + 1, pol),
                 gcc1 = ldg_plane(gc + 2 * ix.g_cstride + wbase + 1, pol);
        e_cur = ~((lo_cur ^ __funnelshift_r(glo0, glo1, sh)) | (hi_cur ^ __funnelshift_r(ghi0, ghi1, sh))) & v_cur &
                __funnelshift_r(gv0, gv1, sh);
        cnt_a = __funnelshift_r(gca0, gca1, sh); cnt_b = __funnelshift_r(gcb0, gcb1, sh); cnt_c = __funnelshift_r(gcc0, gcc1, sh);
        glo0 = glo1; ghi0 = ghi1; gv0 = gv1; gca0 = gca1; gcb0 = gcb1; gcc0 = gcc1;
    }
    int T = 0, c_d = 0;
#pragma unroll 1
    for (int k = 0; k < nch; k++) {
        /* chunk k+1 (the 16-wide run detector needs it across the word boundary) */
        uint32_t nlo, nhi, nv;
        read_chunk(k + 1, &nlo, &nhi, &nv);
        uint32_t e_nxt = 0, na = 0, nb = 0, nc = 0;
        if (seeded && k + 1 < nch) {
            uint32_t glo1 = ldg_plane(ix.g_lo + wbase + k + 2, pol), ghi1 = ldg_plane(ix.g_hi + wbase + k + 2, pol),
                     gv1 = ldg_plane(ix.g_v + wbase + k + 2, pol);
            uint32_t gca1 = ldg_plane(gc + wbase + k + 2, pol), gcb1 = ldg_plane(gc + ix.g_cstride + wbase + k + 2, pol),
                     gcc1 = ldg_plane(gc + 2 * ix.g_cstride + wbase + k + 2, pol);
            e_nxt = ~((nlo ^ __funnelshift_r(glo0, glo1, sh)) | (nhi ^ __funnelshift_r(ghi0, ghi1, sh))) & nv &
                    __funnelshift_r(gv0, gv1, sh);
            na = __funnelshift_r(gca0, gca1, sh); nb = __funnelshift_r(gcb0, gcb1, sh); nc = __funnelshift_r(gcc0, gcc1, sh);
            glo0 = glo1; ghi0 = ghi1; gv0 = gv1; gca0 = gca1; gcb0 = gcb1; gcc0 = gcc1;
        }
        uint32_t m = run16(e_cur, e_nxt) & parity;
        uint32_t c0 = cnt_a & m, c1 = cnt_b & m, c2 = cnt_c & m;
        uint32_t hit = c0 | c1 | c2;
        c_d += __popc(hit);
        T += __popc(c0) + 2 * __popc(c1) + 4 * __popc(c2);
        /* offsets of this chunk the diagonal does not explain -> filter probes, 4 in flight */
        uint32_t om = run16(v_cur, nv) & parity & ~hit;
        while (om) {
            uint32_t key[4];
            unsigned long long w[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ok[u] = om != 0u;
                uint32_t b = ok[u] ? (uint32_t)(__ffs(om) - 1) : 0u;
                om &= om - 1u;
                uint32_t kk = ((__funnelshift_r(hi_cur, nhi, b) & 0xFFFFu) << 16) | (__funnelshift_r(lo_cur, nlo, b) & 0xFFFFu);
                key[u] = rc ? gf_key_revcomp(kk) : kk;
                w[u] = 0;
                if (ok[u]) w[u] = ldg_filter(ix.filter + gf_filter_word(key[u], ix.filter_words), pol);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (ok[u]) T += (int)gf_filter_sites(ix, w[u], key[u], ix.max_sites);
        }
        e_cur = e_nxt; cnt_a = na; cnt_b = nb; cnt_c = nc;
        v_cur = nv; lo_cur = nlo; hi_cur = nhi;
    }
    return T >= need_total && (T - c_d) >= need_minor;
}

template <int W, bool PAIRED>
__global__ void __launch_bounds__(WARPS * 32) k_screen_tpp(ScreenParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t lane = gf_lane(), wib = threadIdx.x >> 5;
    Col c;
    c.base = sm + (size_t)wib * Lay<W>::NWORDS * 32 + lane;
    const GfDevBatch& B = P.b;
    const uint64_t n_warps = (uint64_t)gridDim.x * WARPS;
    const unsigned long long pol = make_policy_keep(), pol_stream = make_policy_stream();
    const uint8_t* const NOBOUND = reinterpret_cast<const uint8_t*>(~(uintptr_t)0);
    const uint8_t* bound1 = B.bytes1 ? B.seq1 + B.bytes1 : NOBOUND;
    const uint8_t* bound2 = (PAIRED && B.bytes2) ? B.seq2 + B.bytes2 : NOBOUND;

    unsigned long long c_seq = 0, c_probes = 0, c_bytes = 0, c_merged = 0;
    uint32_t err = 0;

    for (uint64_t base = ((uint64_t)blockIdx.x * WARPS + wib) * 32; base < B.n; base += n_warps * 32) {
        const uint64_t p = base + lane;
        if (p < B.n) {
            const uint64_t o1 = __ldg(B.s1 + p);
            const int len1 = (int)(__ldg(B.e1 + p) - o1);
            const uint8_t* s1 = B.seq1 + (o1 - B.base1);
            const uint8_t* q1 = B.qual1 + (B.qs1[p] - B.base1);
            int len2 = 0;
            const uint8_t *s2 = nullptr, *q2 = nullptr;
            if (PAIRED) {
                const uint64_t o2 = __ldg(B.s2 + p);
                len2 = (int)(__ldg(B.e2 + p) - o2);
                s2 = B.seq2 + (o2 - B.base2);
                q2 = B.qual2 + (B.qs2[p] - B.base2);
            }
            if (len1 > 32 * W || len2 > 32 * W || len1 < 0 || len2 < 0) {
                err |= 1u;
            } else {
                convert_r1<W>(c, s1, len1, B.seq1, bound1, pol_stream);
                int olen = -1, diff = 0;
                if (PAIRED) {
                    convert_r2_rc<W>(c, s2, len2, B.seq2, bound2, pol_stream);
                    uint32_t r1_bits_unused;
                    olen = find_overlap<W>(c, len1, len2, q1, q2, &diff, &r1_bits_unused);
                }
                const int nseq = olen >= 0 ? 1 : (PAIRED ? 2 : 1);
                for (int sq = 0; sq < nseq; sq++) {
                    int LO = Lay<W>::MLO, HI = Lay<W>::MHI, V = Lay<W>::MV, len;
                    uint32_t meta;
                    if (olen >= 0) {
                        len = build_merged<W>(c, len1, len2, olen, q1, q2);
                        meta = 0u | ((uint32_t)olen << 2) | ((uint32_t)diff << 14);
                        c_merged++;
                    } else if (sq == 0) {
                        LO = Lay<W>::R1LO; HI = Lay<W>::R1HI; V = Lay<W>::R1V;
                        len = len1;
                        meta = 1u;
                    } else {
                        build_r2_forward<W>(c, len2);
                        len = len2;
                        meta = 2u;
                    }
                    c_seq++;
                    c_bytes += (unsigned long long)len;
                    c_probes += (unsigned long long)(len >= 16 ? ((len - 16) >> 1) + 1 : 0);
                    if (screen_sequence(P.ix, c, LO, HI, V, len, P.need_total, P.need_minor, pol)) {
                        uint32_t slot = atomicAdd(&P.counters->n_survivors, 1u);
                        if (slot < P.survivors_cap) P.survivors[slot] = make_uint2((uint32_t)p, meta);
                        else err |= 2u;
                    }
                }
            }
        }
        __syncwarp();
    }
    /* per-thread counts are small (a few hundred pairs per thread): 32-bit warp sums cannot overflow */
    c_seq = __reduce_add_sync(FULL, (unsigned)c_seq);
    c_merged = __reduce_add_sync(FULL, (unsigned)c_merged);
    c_probes = __reduce_add_sync(FULL, (unsigned)c_probes);
    c_bytes = __reduce_add_sync(FULL, (unsigned)c_bytes);
    if (lane == 0) {
        if (c_seq) atomicAdd(&P.counters->n_sequences, c_seq);
        if (c_probes) atomicAdd(&P.counters->n_probes, c_probes);
        if (c_bytes) atomicAdd(&P.counters->seq_bytes, c_bytes);
        if (c_merged) atomicAdd(&P.counters->n_merged, c_merged);
    }
    err = __reduce_or_sync(FULL, err);
    if (err && lane == 0) atomicOr(&P.counters->error_flags, err);
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
