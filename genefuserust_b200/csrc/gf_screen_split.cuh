/*
 * gf_screen_split.cuh — the screen for reads <= 256 bases: thread per pair / per sequence, cut into four kernels so that
 * every warp runs ONE class of work (a fused thread-per-pair kernel averaged 15 of 32 active threads because
 * merged/unmerged, seeded/unseeded, short/long and forward/reverse pairs diverge inside a warp).
 *
 *   k_prep   thread per pair      convert R1 / rc(R2) to bit-planes, fast_merge (read.rs:313-440), write the 1-2 sequences
 *                                 that will be mapped (merged, or R1 and R2) as bit-planes into a sequence store: short
 *                                 sequences from slot 0 upwards, long (merged) ones from the top downwards
 *   k_seed   thread per sequence  8 half-word aligned 16-mers of the first 128 bases -> level-1 filter (L2) -> the first
 *                                 present one -> ONE HBM table lookup; appends the sequence to the seeded list (from the
 *                                 front of one array) or the unseeded list (from its back), entries reserved once per block
 *   k_diag   thread per seeded sequence    compare with the interleaved gene planes along the seed diagonal (exact
 *                                          votes); chunks with offsets the diagonal does not explain are queued per warp
 *                                          and probed in the filter by all lanes together
 *   k_scan   thread per unseeded sequence  level-1 filter probe for the valid even offsets until the outcome is decided
 *
 * Drop rule (indexer.rs:286-360; a k-mer votes at most once per diagonal): with s_i = sites of the k-mer at even offset i,
 * P = #{s_i >= 1}, T = sum min(s_i, 2), c_d = votes on the seed diagonal:  count1 <= P,  count1 + count2 <= T,
 * count2 <= T - c_d.  k_diag drops iff T < need_total or T - c_d < need_minor; k_scan iff P < need_major (T <= 2 P).
 * Survivors re-run the literal algorithm in k_exact, so the screen only has to be conservative.
 *
 * Sequence store layout: slot s, plane word k (lo: 0..NW-1, hi: NW.., valid: 2NW..) at
 * words[((s >> 5) * 3*NW + k) * 32 + (s & 31)] — column-major per group of 32 slots, so a warp reading consecutive slots is
 * coalesced; only words 0 .. ceil(len/32) are written; meta[s] = {pair, source|olen<<2|diff<<14, len, 0}.
 * (A row-major store — one 16-byte {lo, hi, valid, aux} entry per 32 bases, meta and seed folded into the aux words — was
 * measured in round 2: it cuts the DRAM bytes k_diag / k_scan re-read, but every 128-bit load of a warp then touches 32
 * lines instead of 1-3, and k_seed, whose slots ARE consecutive, went from 0.93 to 1.27 ms; the step got 5 % slower.)
 */
#pragma once

namespace split {

using tpp::Col;
using tpp::Lay;

/* The sequence store streams: written once by k_prep, read once by k_seed and once by k_diag or k_scan.  Its loads and stores carry
 * the streaming hint (ld / st.global.cs: first in line for eviction) so that they do not push the L2-resident filter and gene
 * planes out (GF_STORE_STREAMING 0: plain accesses, for comparison). */
#ifndef GF_STORE_STREAMING
#define GF_STORE_STREAMING 1
#endif
template <class T>
__device__ __forceinline__ T st_ld(const T* p) {
#if GF_STORE_STREAMING
    return __ldcs(p);
#else
    return *p;
#endif
}
template <class T>
__device__ __forceinline__ void st_st(T* p, T v) {
#if GF_STORE_STREAMING
    __stcs(p, v);
#else
    *p = v;
#endif
}

struct SeqStore {
    uint32_t* words;
    uint4* meta;
    uint2* seed;              /* per slot: {seed_val, seed_i} */
    uint32_t* list;           /* slot indices: seeded sequences from entry 0 upwards, unseeded ones from entry cap - 1 downwards.
                                 k_seed walks the short slots (at most 32 W bases: every unmerged read) first, then the long
                                 ones, and every block appends once per iteration, so each list keeps the two length classes
                                 apart except in a few warps */
    unsigned int* counters;   /* [0] short slots (from 0 up), [5] long slots (from cap - 1 down), [1] seeded, [2] unseeded entries */
    uint32_t cap;
};
template <int W>
struct SL {
    static constexpr int NW = 2 * W + 1;   /* plane words per sequence (merged reads, + zero pad word) */
    static constexpr int NW3 = 3 * NW;
};
template <int W>
__device__ __forceinline__ uint32_t* slot_words(const SeqStore& st, uint32_t s) {
    return st.words + (size_t)(s >> 5) * SL<W>::NW3 * 32 + (s & 31u);
}

__device__ __forceinline__ uint32_t fs_col(const uint32_t* col, int arr_base, uint32_t bitpos) {
    uint32_t wi = bitpos >> 5;
    return __funnelshift_r(st_ld(&col[(size_t)(arr_base + (int)wi) * 32]), st_ld(&col[(size_t)(arr_base + (int)wi + 1) * 32]), bitpos & 31u);
}

struct PrepParams {
    GfDevBatch b;
    SeqStore st;
    GfMapCounters* counters;
};

/* ---------------------------------------------------------------------------------------------- k_prep */
template <int W, bool PAIRED, bool PACKED>
__global__ void __launch_bounds__(tpp::WARPS * 32, W == 5 ? 8 : 5) k_prep(PrepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t lane = gf_lane(), wib = threadIdx.x >> 5;
    constexpr int PW = Lay<W>::NWORDS; /* R1 + rc(R2) plane arrays */
    Col c;
    c.base = sm + (size_t)wib * PW * 32 + lane;
    const GfDevBatch& B = P.b;
    const uint64_t n_warps = (uint64_t)gridDim.x * tpp::WARPS;
    const unsigned long long pol_stream = tpp::make_policy_normal(); /* measured: evict_first on the read bytes costs ~2 % */
    const uint8_t* const NOBOUND = reinterpret_cast<const uint8_t*>(~(uintptr_t)0);
    const uint8_t* bound1 = B.bytes1 ? B.seq1 + B.bytes1 : NOBOUND;
    const uint8_t* bound2 = (PAIRED && B.bytes2) ? B.seq2 + B.bytes2 : NOBOUND;
    constexpr int NW = SL<W>::NW;
    unsigned c_seq = 0, c_probes = 0, c_bytes = 0, c_merged = 0;
    uint32_t err = 0;

    for (uint64_t base = ((uint64_t)blockIdx.x * tpp::WARPS + wib) * 32; base < B.n; base += n_warps * 32) {
        const uint64_t p = base + lane;
        int len1 = 0, len2 = 0, olen = -1, diff = 0, nseq = 0;
        uint32_t r1_bits = 0; /* which of the (<= 2) overlap mismatches keep the R1 base */
        const uint8_t *q1 = nullptr, *q2 = nullptr;
        if (p < B.n) {
            const uint64_t o1 = __ldg(B.s1 + p), e1 = __ldg(B.e1 + p);
            bool ok = record_ok(o1, e1, B.base1, B.bytes1, 32 * W);
            len1 = (int)(e1 - o1);
            const uint8_t* s1 = B.seq1 + (o1 - B.base1);
            q1 = B.qual1 + (B.qs1[p] - B.base1);
            const uint8_t* s2 = nullptr;
            if (PAIRED) {
                const uint64_t o2 = __ldg(B.s2 + p), e2 = __ldg(B.e2 + p);
                ok = ok && record_ok(o2, e2, B.base2, B.bytes2, 32 * W);
                len2 = (int)(e2 - o2);
                s2 = B.seq2 + (o2 - B.base2);
                q2 = B.qual2 + (B.qs2[p] - B.base2);
            }
            if (!ok) { /* longer than the kernel capacity, or offsets that are not ascending / leave the arena */
                err |= 1u;
                len1 = len2 = 0;
            } else {
                if (PACKED) { /* the host built the planes (gf_pack.cpp) */
                    const uint32_t x1 = __ldg(B.pxo1 + p);
                    tpp::load_r1_packed<W>(c, B.pk1 + __ldg(B.pko1 + p), x1 ? B.pkx1 + (x1 - 1u) : nullptr, len1);
                } else {
                    tpp::convert_r1<W>(c, s1, len1, B.seq1, bound1, pol_stream);
                }
                if (PAIRED) {
                    if (PACKED) {
                        const uint32_t x2 = __ldg(B.pxo2 + p);
                        tpp::load_r2_rc_packed<W>(c, B.pk2 + __ldg(B.pko2 + p), x2 ? B.pkx2 + (x2 - 1u) : nullptr, len2);
                    } else {
                        tpp::convert_r2_rc<W>(c, s2, len2, B.seq2, bound2, pol_stream);
                    }
                    olen = tpp::find_overlap<W>(c, len1, len2, q1, q2, &diff, &r1_bits);
                }
                nseq = olen >= 0 ? 1 : (PAIRED ? 2 : 1);
            }
        }
        /* The first sequence of EVERY pair is "R1[..offset] ++ rc(R2)" (read.rs:369-428): a merged read with offset =
         * len1 - olen, plain R1 with offset = len1 and no rc(R2) part — one code path for merged and unmerged lanes. */
        const bool merged = olen >= 0;
        const int offset = merged ? len1 - olen : len1;
        const int len0 = merged ? offset + len2 : len1;
        /* slots: warp-aggregated allocation (all lanes take part).  Short sequences (<= 32 W bases: every unmerged read)
         * fill the store from slot 0 upwards, long ones (merged reads) from slot cap - 1 downwards, so that the lists
         * k_seed builds refer to nearly contiguous slots and the plane-word loads of k_diag / k_scan stay coalesced. */
        const bool lng = merged && len0 > 32 * W;
        const int mine = lng ? (1 << 16) : nseq;
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, incl, o);
            if ((int)lane >= o) incl += t;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        uint32_t base_s = 0, base_l = 0;
        if (lane == 0) {
            if (total & 0xFFFF) base_s = atomicAdd(&P.st.counters[0], (unsigned)(total & 0xFFFF));
            if (total >> 16) base_l = atomicAdd(&P.st.counters[5], (unsigned)(total >> 16));
        }
        base_s = __shfl_sync(FULL, base_s, 0);
        base_l = __shfl_sync(FULL, base_l, 0);
        const int excl = incl - mine;
        uint32_t slot0 = lng ? P.st.cap - 1u - (base_l + (uint32_t)(excl >> 16)) : base_s + (uint32_t)(excl & 0xFFFF);
        if (nseq && (lng ? base_l + (uint32_t)(excl >> 16) >= P.st.cap : slot0 + (uint32_t)nseq > P.st.cap)) { err |= 2u; nseq = 0; }
        if (nseq) {
            /* positions (<= 2) of the overlap where the R1 base wins: q1 >= '?' and q2 <= '0' (the qualities were looked at in
             * find_overlap: r1_bits) */
            int fix0 = -1, fix1 = -1;
            if (merged && r1_bits) {
                for (int k = 0; 32 * k < olen; k++) {
                    uint32_t mism = tpp::overlap_mism<W>(c, offset, olen, len2, k);
                    while (mism) {
                        int b = __ffs(mism) - 1;
                        mism &= mism - 1;
                        if (r1_bits & 1u) { if (fix0 < 0) fix0 = offset + 32 * k + b; else fix1 = offset + 32 * k + b; }
                        r1_bits >>= 1;
                    }
                }
            }
            uint32_t* w = slot_words<W>(P.st, slot0);
            const int nw = (len0 + 31) >> 5;
            for (int k = 0; k <= nw; k++) { /* words 0 .. ceil(len/32): nothing beyond is ever read */
                uint32_t lo = 0, hi = 0, v = 0;
                if (k < nw) {
                    const int pos0 = 32 * k;
                    if (pos0 < offset) {
                        const uint32_t m = lowmask(offset - pos0);
                        lo = c(Lay<W>::R1LO, k) & m; hi = c(Lay<W>::R1HI, k) & m; v = c(Lay<W>::R1V, k) & m;
                    }
                    const int j0 = pos0 - offset;
                    if (merged && j0 > -32) { lo |= c.win(Lay<W>::C2LO, j0); hi |= c.win(Lay<W>::C2HI, j0); v |= c.win(Lay<W>::C2V, j0); }
                    if (fix0 >= 0) {
                        for (int f = 0; f < 2; f++) {
                            const int fp = f ? fix1 : fix0;
                            if (fp >= 0 && (fp >> 5) == k) {
                                const uint32_t bit = 1u << (fp & 31);
                                lo = (lo & ~bit) | (c(Lay<W>::R1LO, k) & bit);
                                hi = (hi & ~bit) | (c(Lay<W>::R1HI, k) & bit);
                                v = (v & ~bit) | (c(Lay<W>::R1V, k) & bit);
                            }
                        }
                    }
                }
                st_st(&w[(size_t)k * 32], lo); st_st(&w[(size_t)(NW + k) * 32], hi); st_st(&w[(size_t)(2 * NW + k) * 32], v);
            }
            const uint32_t info = merged ? (0u | ((uint32_t)olen << 2) | ((uint32_t)diff << 14)) : 1u;
            st_st(&P.st.meta[slot0], make_uint4((uint32_t)p, info, (uint32_t)len0, 0u));
            c_seq++;
            c_merged += merged ? 1u : 0u;
            c_bytes += (unsigned)len0;
            c_probes += (unsigned)(len0 >= 16 ? ((len0 - 16) >> 1) + 1 : 0);
        }
        if (nseq == 2) {
            /* forward R2 (upper-case validity) from the rc planes */
            uint32_t* w = slot_words<W>(P.st, slot0 + 1u);
            const int nw = (len2 + 31) >> 5;
            for (int k = 0; k <= nw; k++) {
                uint32_t lo = 0, hi = 0, v = 0;
                if (k < nw) {
                    const int pos = len2 - 32 * k - 32;
                    v = __brev(c.win(Lay<W>::VCS, pos));
                    lo = ~__brev(c.win(Lay<W>::C2LO, pos)) & v;
                    hi = __brev(c.win(Lay<W>::C2HI, pos)) & v;
                }
                st_st(&w[(size_t)k * 32], lo); st_st(&w[(size_t)(NW + k) * 32], hi); st_st(&w[(size_t)(2 * NW + k) * 32], v);
            }
            st_st(&P.st.meta[slot0 + 1u], make_uint4((uint32_t)p, 2u, (uint32_t)len2, 0u));
            c_seq++;
            c_bytes += (unsigned)len2;
            c_probes += (unsigned)(len2 >= 16 ? ((len2 - 16) >> 1) + 1 : 0);
        }
        __syncwarp();
    }
    c_seq = __reduce_add_sync(FULL, c_seq);
    c_merged = __reduce_add_sync(FULL, c_merged);
    c_probes = __reduce_add_sync(FULL, c_probes);
    c_bytes = __reduce_add_sync(FULL, c_bytes);
    if (lane == 0) {
        if (c_seq) atomicAdd(&P.counters->n_sequences, (unsigned long long)c_seq);
        if (c_probes) atomicAdd(&P.counters->n_probes, (unsigned long long)c_probes);
        if (c_bytes) atomicAdd(&P.counters->seq_bytes, (unsigned long long)c_bytes);
        if (c_merged) atomicAdd(&P.counters->n_merged, (unsigned long long)c_merged);
    }
    err = __reduce_or_sync(FULL, err);
    if (err && lane == 0) atomicOr(&P.counters->error_flags, err);
}

/* ---------------------------------------------------------------------------------------------- k_seed */
struct SeedParams {
    GfDevIndex ix;
    SeqStore st;
    int need_total, need_minor;
};
template <int W>
__global__ void __launch_bounds__(256) k_seed(SeedParams P) {
    constexpr int NW = SL<W>::NW;
    const uint32_t lane = gf_lane();
    const uint32_t n_short = min(P.st.counters[0], P.st.cap), n_long = min(P.st.counters[5], P.st.cap - n_short);
    const uint32_t n_slots = n_short + n_long; /* virtual index v: short slots, then the long ones in ascending order */
    const unsigned long long pol = make_policy_keep();
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t wib = threadIdx.x >> 5;
    __shared__ uint32_t cls_cnt[8][2], cls_base[8][2];
    /* block-uniform trip count: the list entries are reserved once per block and iteration (below) */
    for (uint32_t b0 = blockIdx.x * blockDim.x; b0 < n_slots; b0 += stride) {
        const uint32_t vi = b0 + threadIdx.x;
        const uint32_t s = vi < n_short ? vi : P.st.cap - n_long + (vi - n_short);
        bool have = vi < n_slots, seeded = false;
        uint32_t seed_val = GF_EMPTY_VAL, seed_i = 0;
        if (have) {
            /* 8 candidate 16-mers at half-word aligned offsets of the first 128 bases (a candidate is one shift + mask of a
             * plane word, two candidates share the three plane loads).  The plane words are requested together with the
             * slot's meta record (one memory round trip less on the critical path); a candidate counts only if it lies inside
             * the read (words beyond the read were never written) and its 16 bases are valid.  Every candidate offset is even
             * and inside the read, i.e. one of pass 1's probe offsets; which ones are tried only decides the diagonal, never
             * the bound.  All go to the L2 filter at once; the HBM table is asked only for candidates the level-1 filter calls
             * present, in read order, and the bucket of the first valid candidate is prefetched into L2 meanwhile. */
            const uint32_t* col = slot_words<W>(P.st, s);
            uint32_t plo[4], phi[4], pv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                plo[j] = st_ld(&col[(size_t)j * 32]); phi[j] = st_ld(&col[(size_t)(NW + j) * 32]); pv[j] = st_ld(&col[(size_t)(2 * NW + j) * 32]);
            }
            const uint4 m = st_ld(&P.st.meta[s]);
            const int len = (int)m.z;
            const int nprobe = len >= 16 ? ((len - 16) >> 1) + 1 : 0;
            if (nprobe == 0 && P.need_total > 0) have = false; /* cannot reach the gate: dropped here */
            if (have && nprobe > 0) {
                uint32_t key[8], okm = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    key[2 * j] = (phi[j] << 16) | (plo[j] & 0xFFFFu);
                    key[2 * j + 1] = (phi[j] & 0xFFFF0000u) | (plo[j] >> 16);
                    if (32 * j + 16 <= len && (pv[j] & 0xFFFFu) == 0xFFFFu) okm |= 1u << (2 * j);
                    if (32 * j + 32 <= len && (pv[j] >> 16) == 0xFFFFu) okm |= 2u << (2 * j);
                }
                if (okm) {
                    const uint32_t t0 = (uint32_t)__ffs(okm) - 1u;
                    uint32_t k0 = key[0];
#pragma unroll
                    for (int t = 1; t < 8; t++) if (t0 == (uint32_t)t) k0 = key[t];
                    tpp::prefetch_l2(P.ix.table + 2ull * gf_home_bucket(k0, P.ix.bucket_shift));
                }
                unsigned long long fw[8];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    fw[t] = 0;
                    if ((okm >> t) & 1u) fw[t] = ldg_filter(P.ix.filter + gf_filter_word(key[t], P.ix.filter_words), pol);
                }
                /* candidates the main filter calls present (the NORMAL / unique distinction is left to the table: asking
                 * the second-level filter would cost another dependent L2 gather per candidate) */
                uint32_t pm = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    uint32_t al, ah;
                    gf_filter_masks(key[t], &al, &ah);
                    const uint32_t wl = (uint32_t)fw[t], wh = (uint32_t)(fw[t] >> 32);
                    if (((okm >> t) & 1u) && (wl & al) == al && (wh & ah) == ah) pm |= 1u << t;
                }
                /* an unseeded sequence goes to k_scan, which need not probe these candidates again: it starts from their
                 * count (bits 0-7 of seed.y) and skips the offsets (bits 8-15: which were valid) */
                seed_i = pm | (okm << 8);
                while (pm) { /* usually one iteration: the first present candidate of an on-target read is unique */
                    const int t = __ffs(pm) - 1;
                    pm &= pm - 1u;
                    uint32_t kt = key[0];
#pragma unroll
                    for (int u = 1; u < 8; u++) if (t == u) kt = key[u];
                    const uint32_t val = gf_table_find(P.ix, kt);
                    if (val != GF_EMPTY_VAL && (val >> 30) == GF_KIND_UNIQUE) {
                        seed_val = val;
                        seed_i = 16u * (uint32_t)t;
                        seeded = true;
                        break;
                    }
                }
            }
            if (have) st_st(&P.st.seed[s], make_uint2(seed_val, seed_i));
        }
        /* the two lists, aggregated over the block: one atomic per class, block and iteration (the list counters are
         * single hot addresses; a per-warp atomic on them serialises in L2) */
        const int cls = have ? (seeded ? 0 : 1) : -1;
        const uint32_t m0 = __ballot_sync(FULL, cls == 0), m1 = __ballot_sync(FULL, cls == 1);
        if (lane == 0) { cls_cnt[wib][0] = (uint32_t)__popc(m0); cls_cnt[wib][1] = (uint32_t)__popc(m1); }
        __syncthreads();
        if (threadIdx.x < 2) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += cls_cnt[w][threadIdx.x];
            uint32_t bb = tot ? atomicAdd(&P.st.counters[1 + threadIdx.x], tot) : 0u;
#pragma unroll
            for (int w = 0; w < 8; w++) { cls_base[w][threadIdx.x] = bb; bb += cls_cnt[w][threadIdx.x]; }
        }
        __syncthreads();
        if (cls == 0) st_st(&P.st.list[cls_base[wib][0] + __popc(m0 & gf_lanemask_lt())], s);
        else if (cls == 1) st_st(&P.st.list[P.st.cap - 1u - (cls_base[wib][1] + __popc(m1 & gf_lanemask_lt()))], s);
    }
}

/* ---------------------------------------------------------------------------------------------- k_diag / k_scan */
struct ClassParams {
    GfDevIndex ix;
    SeqStore st;
    uint2* survivors;
    uint32_t survivors_cap;
    GfMapCounters* counters;
    int need_total, need_minor;
};
__device__ __forceinline__ void push_survivor(const ClassParams& P, const uint4& m) {
    uint32_t slot = atomicAdd(&P.counters->n_survivors, 1u);
    if (slot < P.survivors_cap) P.survivors[slot] = make_uint2(m.x, m.y);
    else atomicOr(&P.counters->error_flags, 2u);
}
/* unseeded sequences: the valid even offsets get a filter probe until the outcome is decided.
 * Bound (indexer.rs:286-360): a k-mer votes at most ONCE for any one diagonal (the sites of a key are distinct), so with
 * s_i = sites of the k-mer at offset i:  count1 <= P = #{i : s_i >= 1}  and  count1 + count2 <= T = sum min(s_i, 2).
 * The gate needs count1 >= need_major and count2 >= need_minor, hence P >= need_major and T >= need_major + need_minor.
 * Only the level-1 filter is asked here (present / absent), i.e. T is bounded by 2 P and the test is P >= need_major (and
 * 2 P >= need_major + need_minor); probing stops as soon as the offsets that are left cannot lift P over the threshold
 * (off-target reads: after ~3/4 of their offsets). */
template <int W>
__global__ void __launch_bounds__(256) k_scan(ClassParams P) {
    constexpr int NW = SL<W>::NW;
    const unsigned long long pol = make_policy_keep();
    const GfDevIndex& ix = P.ix;
    const int need_major = P.need_total - P.need_minor;
    const uint32_t n = P.st.counters[2];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t s = st_ld(&P.st.list[P.st.cap - 1u - t]);
        const uint4 m = st_ld(&P.st.meta[s]);
        const int len = (int)m.z, nch = (len + 31) >> 5;
        const uint32_t* col = slot_words<W>(P.st, s);
        uint32_t lo = st_ld(&col[0]), hi = st_ld(&col[(size_t)NW * 32]), v = st_ld(&col[(size_t)2 * NW * 32]);
        /* k_seed already asked the filter about the 16-mers at offsets 0, 16, ..., 112 (seed.y: bits 0-7 present,
         * bits 8-15 valid = probed) */
        const uint32_t pre = st_ld(&P.st.seed[s]).y, pre_valid = (pre >> 8) & 0xFFu;
        int Pn = __popc(pre & 0xFFu);
        bool dead = false;
#pragma unroll 1
        for (int k = 0; k < nch && !dead; k++) {
            const uint32_t nlo = st_ld(&col[(size_t)(k + 1) * 32]), nhi = st_ld(&col[(size_t)(NW + k + 1) * 32]), nv = st_ld(&col[(size_t)(2 * NW + k + 1) * 32]);
            uint32_t om = run16(v, nv) & 0x55555555u;
            if (k < 4) om &= ~0x00010001u;                          /* offsets 32 k and 32 k + 16: probed by k_seed */
            const int beyond = len - 16 - 32 * (k + 1);            /* last probe offset relative to the next chunk */
            const int rem_after = (beyond >= 0 ? (beyond >> 1) + 1 : 0) - (k < 3 ? __popc(pre_valid >> (2 * (k + 1))) : 0);
            while (om) {
                uint32_t key[4];
                unsigned long long w[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    ok[u] = om != 0u;
                    const uint32_t b = ok[u] ? (uint32_t)(__ffs(om) - 1) : 0u;
                    om &= om - 1u;
                    key[u] = ((__funnelshift_r(hi, nhi, b) & 0xFFFFu) << 16) | (__funnelshift_r(lo, nlo, b) & 0xFFFFu);
                    w[u] = 0;
                    if (ok[u]) w[u] = ldg_filter(ix.filter + gf_filter_word(key[u], ix.filter_words), pol);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (ok[u] && gf_filter_present(w[u], key[u])) Pn++;
                const int rem = __popc(om) + rem_after;
                if (P.need_total > 0 && Pn + rem < need_major) { dead = true; break; }
            }
            lo = nlo; hi = nhi; v = nv;
        }
        /* T <= 2 Pn without the unique / dupe distinction: the T conditions follow from Pn >= need_major when
         * need_minor <= need_major (checked by the host; otherwise 2 Pn is compared) */
        if (P.need_total <= 0 || (!dead && Pn >= need_major && 2 * Pn >= P.need_total)) push_survivor(P, m);
    }
}

/* one interleaved gene-plane entry {lo, hi, valid, count bit 0, 1, 2, -, -}: a single 256-bit L2 load, kept in L2 */
struct GeneWord { uint32_t lo, hi, v, a, b, c, pad0, pad1; };
__device__ __forceinline__ GeneWord ldg_gene_word(const uint32_t* base, uint32_t w, unsigned long long pol) {
    GeneWord g;
    asm volatile("ld.global.nc.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(g.lo), "=r"(g.hi), "=r"(g.v), "=r"(g.a), "=r"(g.b), "=r"(g.c), "=r"(g.pad0), "=r"(g.pad1)
                 : "l"(base + 8ull * w), "l"(pol));
    return g;
}
constexpr int DIAG_Q = 128; /* queue entries per warp; flushed above DIAG_Q - 32 (a chunk adds <= 32) */
template <int W>
__global__ void __launch_bounds__(256, 5) k_diag(ClassParams P) {
    constexpr int NW = SL<W>::NW;
    __shared__ uint32_t q_lo0[8][DIAG_Q], q_lo1[8][DIAG_Q], q_hi0[8][DIAG_Q], q_hi1[8][DIAG_Q], q_om[8][DIAG_Q], q_meta[8][DIAG_Q];
    __shared__ int t_sh[8][32];
    __shared__ unsigned q_cnt[8];
    const uint32_t lane = gf_lane(), wib = threadIdx.x >> 5;
    const unsigned long long pol = make_policy_keep();
    const GfDevIndex& ix = P.ix;
    auto flush = [&]() {
        __syncwarp();
        const int total = (int)q_cnt[wib];
#pragma unroll 2
        for (int base = 0; base < total; base += 2) {
            const int e = base + (int)(lane >> 4);
            if (e < total) {
                const uint32_t om = q_om[wib][e], meta = q_meta[wib][e];
                const uint32_t b = 2u * (lane & 15u) + ((meta >> 9) & 1u);
                if ((om >> b) & 1u) {
                    const uint32_t kk = ((__funnelshift_r(q_hi0[wib][e], q_hi1[wib][e], b) & 0xFFFFu) << 16) |
                                        (__funnelshift_r(q_lo0[wib][e], q_lo1[wib][e], b) & 0xFFFFu);
                    const uint32_t key = ((meta >> 8) & 1u) ? gf_key_revcomp(kk) : kk;
                    if (gf_filter_present(ldg_filter(ix.filter + gf_filter_word(key, ix.filter_words), pol), key))
                        atomicAdd(&t_sh[wib][meta & 31u], 2); /* min(sites, 2) <= 2 */
                }
            }
        }
        __syncwarp();
        if (lane == 0) q_cnt[wib] = 0;
        __syncwarp();
    };
    const uint32_t n = P.st.counters[1];
    for (uint32_t t0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); t0 < n; t0 += gridDim.x * blockDim.x) {
        const bool have = t0 + lane < n;
        const uint32_t s = have ? st_ld(&P.st.list[t0 + lane]) : 0u;
        const uint4 m = have ? st_ld(&P.st.meta[s]) : make_uint4(0, 0, 0, 0);
        const uint2 sd = have ? st_ld(&P.st.seed[s]) : make_uint2(0, 0);
        const int len = (int)m.z, nch = have ? (len + 31) >> 5 : 0;
        const uint32_t* col = slot_words<W>(P.st, s);
        const bool rc = (sd.x & GF_SITE_STRAND) != 0;
        const uint32_t goff = sd.x & GF_SITE_GOFF_MASK;
        const uint32_t D = rc ? goff + sd.y - (uint32_t)len + 1u : goff - sd.y;
        const uint32_t parity = (rc && (len & 1)) ? 0xAAAAAAAAu : 0x55555555u;
        const uint32_t wbase = D >> 5, sh = D & 31u;
        const uint32_t* gi = rc ? ix.g_ir : ix.g_if;
        t_sh[wib][lane] = 0;
        if (lane == 0) q_cnt[wib] = 0;
        __syncwarp();
        /* read chunk k in the orientation of the comparison */
        auto read_chunk = [&](int k, uint32_t* lo, uint32_t* hi, uint32_t* v) {
            if (k >= nch) { *lo = *hi = *v = 0; return; }
            if (!rc) { *lo = st_ld(&col[(size_t)k * 32]); *hi = st_ld(&col[(size_t)(NW + k) * 32]); *v = st_ld(&col[(size_t)(2 * NW + k) * 32]); return; }
            int pos = len - 32 * k - 32;
            uint32_t a, b, cc;
            if (pos >= 0) { a = fs_col(col, 0, (uint32_t)pos); b = fs_col(col, NW, (uint32_t)pos); cc = fs_col(col, 2 * NW, (uint32_t)pos); }
            else { a = st_ld(&col[0]) << (-pos); b = st_ld(&col[(size_t)NW * 32]) << (-pos); cc = st_ld(&col[(size_t)2 * NW * 32]) << (-pos); }
            uint32_t vv = __brev(cc);
            *v = vv;
            *lo = ~__brev(a) & vv;
            *hi = __brev(b);
        };
        GeneWord g0 = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t lo_cur = 0, hi_cur = 0, v_cur = 0, e_cur = 0, cnt_a = 0, cnt_b = 0, cnt_c = 0;
        if (have) {
            g0 = ldg_gene_word(gi, wbase, pol);
            const GeneWord g1 = ldg_gene_word(gi, wbase + 1, pol);
            read_chunk(0, &lo_cur, &hi_cur, &v_cur);
            e_cur = ~((lo_cur ^ __funnelshift_r(g0.lo, g1.lo, sh)) | (hi_cur ^ __funnelshift_r(g0.hi, g1.hi, sh))) & v_cur &
                    __funnelshift_r(g0.v, g1.v, sh);
            cnt_a = __funnelshift_r(g0.a, g1.a, sh); cnt_b = __funnelshift_r(g0.b, g1.b, sh); cnt_c = __funnelshift_r(g0.c, g1.c, sh);
            g0 = g1;
        }
        int T = 0, c_d = 0;
        const int max_nch = (int)__reduce_max_sync(FULL, (unsigned)nch);
#pragma unroll 1
        for (int k = 0; k < max_nch; k++) {
            if (k < nch) {
                uint32_t nlo, nhi, nv;
                read_chunk(k + 1, &nlo, &nhi, &nv);
                uint32_t e_nxt = 0, na = 0, nb = 0, nc = 0;
                if (k + 1 < nch) {
                    const GeneWord g1 = ldg_gene_word(gi, wbase + k + 2, pol);
                    e_nxt = ~((nlo ^ __funnelshift_r(g0.lo, g1.lo, sh)) | (nhi ^ __funnelshift_r(g0.hi, g1.hi, sh))) & nv &
                            __funnelshift_r(g0.v, g1.v, sh);
                    na = __funnelshift_r(g0.a, g1.a, sh); nb = __funnelshift_r(g0.b, g1.b, sh); nc = __funnelshift_r(g0.c, g1.c, sh);
                    g0 = g1;
                }
                uint32_t mm = run16(e_cur, e_nxt) & parity;
                uint32_t c0 = cnt_a & mm, c1 = cnt_b & mm, c2 = cnt_c & mm;
                uint32_t hit = c0 | c1 | c2;
                c_d += __popc(hit);
                T += __popc(hit) + __popc(c1 | c2); /* min(sites, 2) per offset: a k-mer votes once per diagonal */
                uint32_t om = run16(v_cur, nv) & parity & ~hit;
                e_cur = e_nxt; cnt_a = na; cnt_b = nb; cnt_c = nc;
                /* queue this chunk's unexplained offsets (one entry) */
                if (om) {
                    const unsigned pos = atomicAdd(&q_cnt[wib], 1u);
                    q_lo0[wib][pos] = lo_cur; q_lo1[wib][pos] = nlo; q_hi0[wib][pos] = hi_cur; q_hi1[wib][pos] = nhi;
                    q_om[wib][pos] = om;
                    q_meta[wib][pos] = lane | (rc ? 0x100u : 0u) | ((parity & 1u) ? 0u : 0x200u);
                }
                v_cur = nv; lo_cur = nlo; hi_cur = nhi;
            }
            __syncwarp();
            if (q_cnt[wib] > (unsigned)(DIAG_Q - 32)) flush();
        }
        flush();
        T += t_sh[wib][lane];
        if (have && (P.need_total <= 0 || P.need_minor <= 0 || (T >= P.need_total && (T - c_d) >= P.need_minor))) push_survivor(P, m);
        __syncwarp();
    }
}

}  // namespace split
