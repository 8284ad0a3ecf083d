/*
 * gf_matcher.cu — the Matcher pass of FusionMapper::remove_alignables (/root/reference/src/core/fusion_mapper.rs:488-542,
 * src/core/matcher.rs) as ONE streaming scan of the reference on the GPU.  include/genefuse_gpu.h says what the (degenerate)
 * reference code computes; here is how it maps to kernels.
 *
 *   k_ref_scan     HBM streaming, 1 byte per base read, nothing written.  Two aligned 32-byte sectors of reference text per
 *                  thread; a warp covers 63 consecutive sectors + the sector before them (context only), so neighbouring
 *                  plane words travel by shuffle: no shared memory, no barrier, one launch per resident reference (a warp
 *                  finds its segment from a small prefix table).  32 ASCII bases -> two plane words (valid = ACGT in either
 *                  case = to_ascii_uppercase, matcher.rs:143-148; isA) with ONE branch-free SWAR routine in the nibble domain
 *                  (sector_va), then the keep rule of index_contig_bytes (:227-289) for all 32 positions at once:
 *                      position i is kept  <=>  base i is ACGT, i < len - 16, and walking back from i-1 over at most 15
 *                      positions the first base that is not 'A' is a non-ACGT byte (or lies before the contig start), or
 *                      all 15 are 'A'                      (<=> the rolling 32-bit value of the run is < 4)
 *                  and its key is the base's own 2-bit code.  Almost every sector takes the short way out: all 47 bytes in
 *                  reach are valid and no aligned byte of the isA plane is 0xFF (a run of 15 'A's always covers one), or no
 *                  byte is valid at all — then nothing is kept.  Otherwise the rule is evaluated as a carry look-ahead over a
 *                  15-base window and the few kept bases are fetched again for their code.
 *   k_ref_scan<EMIT> second pass, only when some key ended with 1..50 positions (never on a real genome): the same scan,
 *                  appending (contig, position) of those keys to a small list (map_to_index votes with them, :426-433).
 *   k_seq_present  warp per surviving read: which base codes occur at a k-mer start of the read (upper case only,
 *                  make_kmer :849-885) and of its reverse complement (case-insensitive, sequence.rs:52-60) = what
 *                  init_bloom_filter (:63-88) sets, and what map_to_index looks up.
 *   k_seq_decide   thread per read: the panic pre-condition of map_to_index (:388-529) for both orientations.
 */
#include <algorithm>
#include <cstring>
#include <vector>

#include "gf_internal.h"
#include "gf_swar.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WT_SECTORS = 63;          /* counted sectors per warp tile: two per lane, lane 0's first is the sector before them */
constexpr uint64_t RS_CHUNK = 64ull << 20; /* staging buffer bytes (x2, double buffered) */
constexpr int RS_LIST_CAP = 64;            /* positions kept per key (votes need at most 50) */

struct RefSeg {      /* a run of bases of one contig, resident at device address p */
    const uint8_t* p;    /* first COUNTED base */
    uint64_t n;          /* counted bases */
    uint64_t pos0;       /* contig position of p[0] */
    uint64_t contig_len;
    uint32_t lead;       /* readable context bytes before p (>= 16 unless pos0 == 0) */
    uint32_t contig;
    uint64_t wt0;        /* index of the segment's first warp tile in the launch (segments are in ascending wt0 order) */
    uint64_t n_wt;       /* warp tiles: ceil(sectors / 31) */
};
struct RefScanOut {
    unsigned long long count[4];
    unsigned int n_listed[4];
    unsigned int list_contig[4][RS_LIST_CAP];
    unsigned int list_pos[4][RS_LIST_CAP];
};

/* four bytes of a sector that straddles an end of the segment's readable bytes (first / last sector of a segment only):
 * byte by byte, bytes outside read as 0 = invalid; kept out of line */
__device__ __noinline__ uint32_t load_word_slow(const uint8_t* q, const uint8_t* lo, const uint8_t* hi) {
    uint32_t w = 0;
#pragma unroll
    for (int t = 0; t < 4; t++)
        if (q + t >= lo && q + t < hi) w |= (uint32_t)__ldg(q + t) << (8 * t);
    return w;
}
__device__ __forceinline__ void load_sector(const uint8_t* sec, const uint8_t* lo, const uint8_t* hi, uint4* a, uint4* b) {
    if (sec >= lo && sec + 32 <= hi) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w)
                     : "l"(sec));
        return;
    }
    *a = make_uint4(load_word_slow(sec, lo, hi), load_word_slow(sec + 4, lo, hi), load_word_slow(sec + 8, lo, hi), load_word_slow(sec + 12, lo, hi));
    *b = make_uint4(load_word_slow(sec + 16, lo, hi), load_word_slow(sec + 20, lo, hi), load_word_slow(sec + 24, lo, hi), load_word_slow(sec + 28, lo, hi));
}
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) { /* (a & m) | (b & ~m) as ONE logic operation */
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}
/* 8 bases (two words) -> bit 3 of every nibble of *v / *pa: valid (ACGT, either case) / 'A' or 'a'.  z holds the low nibbles
 * (nibble 2 i = byte i of w0, nibble 2 i + 1 = byte i of w1), y the high nibbles in the same order; with L / H the nibbles of a
 * byte:  valid <=> H = 01x?, L3 = 0, H0 = L2 & ~L1 (only 'T' = 0x54 has the 0x10 bit), L0 = ~H0
 * (A 0x41, C 0x43, G 0x47 -> L = 0001, 0011, 0111 with H0 = 0; T -> L = 0100 with H0 = 1); 'A' <=> valid, L2 = L1 = 0.
 * The bits are lined up at the TOP of the nibble because that takes left shifts only, which are multiplies on the FMA pipe: the
 * kernel is bound by the ALU pipe (logic / shift / compare, one warp instruction per 2 cycles and scheduler), not by issue.
 * tests/test_swar_bits.py models this routine on the CPU for every byte value with the constants read from THIS file. */
__device__ __forceinline__ void group_va(uint32_t w0, uint32_t w1, uint32_t* v, uint32_t* pa) {
    const uint32_t z = bitsel(w0, w1 << 4, 0x0F0F0F0Fu);
    const uint32_t y = bitsel(w0 >> 4, w1, 0x0F0F0F0Fu);
    const uint32_t l2 = z << 1, l1 = z << 2, l0 = z << 3, h2 = y << 1, h0 = y << 3; /* L2, L1, L0, H2, H0 at bit 3 */
    const uint32_t t1 = ~y & h2 & ~z;
    const uint32_t t2 = ~(h0 ^ (l2 & ~l1));
    const uint32_t t3 = t1 & t2 & 0x88888888u;
    const uint32_t vv = t3 & (l0 ^ h0);
    *v = vv;
    *pa = vv & ~l2 & ~l1;
}
constexpr uint32_t VA_ALL = 0x88888888u; /* a group_va word with all 8 bases set */
/* 32 bases -> plane words (bit t = byte t of the sector); bytes outside the segment were read as 0 = invalid.  A nibble's
 * bit 3 is moved to its base's place in the top byte by one multiply (the same partial-product trick as swar::block16). */
__device__ __forceinline__ void sector_va(const uint4& a, const uint4& b, uint32_t* v, uint32_t* pa) {
    uint32_t v0, v1, v2, v3, a0, a1, a2, a3;
    group_va(a.x, a.y, &v0, &a0);
    group_va(a.z, a.w, &v1, &a1);
    group_va(b.x, b.y, &v2, &a2);
    group_va(b.z, b.w, &v3, &a3);
    constexpr uint32_t M = 0x00204081u;
    *v = __byte_perm(__byte_perm(v0 * M, v1 * M, 0x0073u), __byte_perm(v2 * M, v3 * M, 0x0073u), 0x5410u);
    *pa = __byte_perm(__byte_perm(a0 * M, a1 * M, 0x0073u), __byte_perm(a2 * M, a3 * M, 0x0073u), 0x5410u);
}

/* the keep rule of one sector in full, from memory (rare: a sector with an invalid byte in reach or with 8 aligned 'A's in
 * reach); returns the number of kept positions per key, one byte each (A, T, C, G) */
template <bool EMIT>
__device__ __noinline__ uint32_t keep_rule(const uint8_t* sec, const RefSeg* __restrict__ sgp, RefScanOut* __restrict__ out, uint32_t emit_mask) {
    const RefSeg sg = *sgp;
    const uint8_t *lo_b = sg.p - sg.lead, *hi_b = sg.p + sg.n;
    uint4 a, b;
    uint32_t v, pa, v_prev, pa_prev;
    load_sector(sec, lo_b, hi_b, &a, &b);
    sector_va(a, b, &v, &pa);
    load_sector(sec - 32, lo_b, hi_b, &a, &b);
    sector_va(a, b, &v_prev, &pa_prev);
    uint32_t cnt = 0;
    const unsigned long long v64 = ((unsigned long long)v << 32) | v_prev;
    /* blocked[i] <=> some valid non-'A' base (a generate G) lies d <= 15 positions behind i with nothing but 'A's (propagate
     * P) in between: a carry look-ahead over a window of 15, in 3 doubling steps + 3 combines (bit b of X << d = position
     * i - d).  Non-ACGT bytes neither generate nor propagate, so a run start is never blocked; 15 'A's are a window without
     * a generate.  kept = valid & ~blocked. */
    const unsigned long long P1 = ((unsigned long long)pa << 32) | pa_prev, G1 = v64 & ~P1;
    const unsigned long long G2 = G1 | (P1 & (G1 << 1)), P2 = P1 & (P1 << 1);       /* window of 2 positions ending at j */
    const unsigned long long G4 = G2 | (P2 & (G2 << 2)), P4 = P2 & (P2 << 2);       /* 4 */
    const unsigned long long G8 = G4 | (P4 & (G4 << 4)), P8 = P4 & (P4 << 4);       /* 8 */
    /* window of 15 = 8 + 4 + 2 + 1 positions ending at j */
    unsigned long long G15 = G8 | (P8 & (G4 << 8));
    const unsigned long long P12 = P8 & (P4 << 8);
    G15 |= P12 & (G2 << 12);
    const unsigned long long P14 = P12 & (P2 << 12);
    G15 |= P14 & (G1 << 14);
    const unsigned long long ok = ~(G15 << 1); /* blocked[i] = window ending at i - 1 holds a connected generate */
    /* counted positions of this sector: inside [p, p + n) and i < contig_len - 16 (`0..len-16`, :237-243) */
    const long long first = (long long)(sec - sg.p); /* index of byte 0 of the sector relative to p */
    long long lim = (long long)sg.n;
    if (sg.contig_len >= 16) lim = min(lim, (long long)(sg.contig_len - 16) - (long long)sg.pos0);
    else lim = 0;
    uint32_t cm = 0;
    {
        const long long b0 = max(0ll, -first), b1 = min(32ll, lim - first); /* bits [b0, b1) */
        if (b1 > b0) cm = (b1 >= 32 ? 0xFFFFFFFFu : ((1u << (int)b1) - 1u)) & ~((1u << (int)b0) - 1u);
    }
    uint32_t kept = (uint32_t)((v64 & ok) >> 32) & cm;
    while (kept) { /* a run start or a base behind 15 'A's; its key is its own code (A 0, T 1, C 2, G 3) */
        const int bit = __ffs(kept) - 1;
        kept &= kept - 1;
        const uint32_t ch = __ldg(sec + bit);
        const uint32_t k = ((ch >> 2) & 1u) | (ch & 2u);
        cnt += 1u << (8 * k);
        if (EMIT && ((emit_mask >> k) & 1u)) {
            const unsigned slot = atomicAdd(&out->n_listed[k], 1u);
            if (slot < (unsigned)RS_LIST_CAP) {
                out->list_contig[k][slot] = sg.contig;
                out->list_pos[k][slot] = (unsigned int)(sg.pos0 + (uint64_t)(first + bit));
            }
        }
    }
    return cnt;
}

template <bool EMIT>
__global__ void __launch_bounds__(RS_THREADS, 4) k_ref_scan(const RefSeg* __restrict__ segs, uint32_t n_segs, uint64_t total_wt,
                                                            RefScanOut* __restrict__ out, uint32_t emit_mask) {
    __shared__ unsigned int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t stride = (uint64_t)(32 * RS_WT_SECTORS) * n_warps;
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    /* the current segment: first / one-past-last readable byte and its tile range [wt_lo, wt_end); lane l's sector of warp tile
     * g is `sec`, advanced by `stride` per step and recomputed when the warp enters another segment (tiles are visited in
     * ascending order: only ever forward).  Every sector of a tile strictly inside (wt_lo, wt_end - 1) is readable as a whole. */
    uint32_t si = 0;
    const uint8_t *lo_b = nullptr, *hi_b = nullptr, *sec = nullptr;
    uint64_t wt_lo = 0, wt_end = 0;
    uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    auto enter = [&]() {
        RefSeg sg = segs[si];
        while (g >= sg.wt0 + sg.n_wt && si + 1 < n_segs) sg = segs[++si]; /* warp-uniform */
        lo_b = sg.p - sg.lead;
        hi_b = sg.p + sg.n;
        wt_lo = sg.wt0;
        wt_end = sg.wt0 + sg.n_wt;
        sec = reinterpret_cast<const uint8_t*>((uintptr_t)sg.p & ~(uintptr_t)31) + 32ll * (2 * (long long)lane - 1) +
              (long long)(32 * RS_WT_SECTORS) * (long long)(g - sg.wt0);
    };
    if (g < total_wt) enter();
#pragma unroll 1
    for (; g < total_wt; g += n_warps, sec += stride) {
        if (g >= wt_end) enter();
        /* two consecutive sectors per lane (the loop's own bookkeeping is paid once for 64 bytes) */
        uint4 a, b, c, d;
        if (g > wt_lo && g + 1 < wt_end) {
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                         : "l"(sec));
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w), "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)
                         : "l"(sec + 32));
        } else {
            load_sector(sec, lo_b, hi_b, &a, &b);
            load_sector(sec + 32, lo_b, hi_b, &c, &d);
        }
        /* Everything the short way out needs is visible in the nibble domain (no gathers): nothing is kept in the context
         * sector; in a sector without a valid base; and in a sector where all 32 bases and the 16 before them are valid and
         * none of the aligned groups of 8 from position -8 to 31 is all 'A' (a run of 15 'A's ending at -1 .. 30 covers one
         * of the groups -8 .. 23): every position then has a valid non-'A' base behind it within 15. */
        uint32_t v0, v1, v2, v3, a0, a1, a2, a3;
        group_va(a.x, a.y, &v0, &a0);
        group_va(a.z, a.w, &v1, &a1);
        group_va(b.x, b.y, &v2, &a2);
        group_va(b.z, b.w, &v3, &a3);
        const bool tail1 = (v2 & v3) == VA_ALL && a3 != VA_ALL; /* what the sector after this one needs of it */
        const bool inner1 = (v0 & v1 & v2 & v3) == VA_ALL && a0 != VA_ALL && a1 != VA_ALL && a2 != VA_ALL && a3 != VA_ALL;
        const bool some1 = (v0 | v1 | v2 | v3) != 0u;
        group_va(c.x, c.y, &v0, &a0);
        group_va(c.z, c.w, &v1, &a1);
        group_va(d.x, d.y, &v2, &a2);
        group_va(d.z, d.w, &v3, &a3);
        const bool tail2 = (v2 & v3) == VA_ALL && a3 != VA_ALL;
        const bool inner2 = (v0 & v1 & v2 & v3) == VA_ALL && a0 != VA_ALL && a1 != VA_ALL && a2 != VA_ALL && a3 != VA_ALL;
        const bool some2 = (v0 | v1 | v2 | v3) != 0u;
        const bool prev_tail = __shfl_up_sync(0xFFFFFFFFu, (int)tail2, 1) != 0; /* the lane before holds the two sectors before */
        if (lane != 0 && some1 && !(prev_tail && inner1)) {
            const uint32_t kc = keep_rule<EMIT>(sec, segs + si, out, emit_mask);
            c0 += kc & 0xFFu; c1 += (kc >> 8) & 0xFFu; c2 += (kc >> 16) & 0xFFu; c3 += kc >> 24;
        }
        if (some2 && !(tail1 && inner2)) {
            const uint32_t kc = keep_rule<EMIT>(sec + 32, segs + si, out, emit_mask);
            c0 += kc & 0xFFu; c1 += (kc >> 8) & 0xFFu; c2 += (kc >> 16) & 0xFFu; c3 += kc >> 24;
        }
    }
    c0 = __reduce_add_sync(0xFFFFFFFFu, c0); c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
    c2 = __reduce_add_sync(0xFFFFFFFFu, c2); c3 = __reduce_add_sync(0xFFFFFFFFu, c3);
    __syncthreads();
    if (lane == 0) {
        if (c0) atomicAdd(&s_cnt[0], c0);
        if (c1) atomicAdd(&s_cnt[1], c1);
        if (c2) atomicAdd(&s_cnt[2], c2);
        if (c3) atomicAdd(&s_cnt[3], c3);
    }
    __syncthreads();
    if (!EMIT && threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&out->count[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

/* per read: bits 0..3 = base codes at the k-mer starts [0, len-16] of the read (upper-case ACGT only), bits 4..7 = the same
 * for its reverse complement (= complement codes of bytes [15, len-1], either case), bit 8 = shorter than 15 bases */
__global__ void k_seq_present(const uint8_t* __restrict__ seqs, const unsigned long long* __restrict__ off, uint64_t n,
                              uint16_t* __restrict__ present, unsigned int* __restrict__ bloom) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const uint64_t a = off[w], e = off[w + 1];
    const long long len = (long long)(e - a);
    uint32_t pf = 0, pr = 0;
    for (long long j = lane; j < len; j += 32) {
        const uint32_t c = seqs[a + (uint64_t)j];
        if (j <= len - 16 && gf_is_acgt_upper(c)) pf |= 1u << ((gf_code_hi(c) << 1) | gf_code_lo(c));
        if (j >= 15 && gf_is_acgt_upper(c & 0xDFu)) pr |= 1u << (((gf_code_hi(c) << 1) | gf_code_lo(c)) ^ 1u); /* complement = code ^ 1 */
    }
    pf = __reduce_or_sync(0xFFFFFFFFu, pf);
    pr = __reduce_or_sync(0xFFFFFFFFu, pr);
    if (lane == 0) {
        const uint32_t p = pf | (pr << 4) | (len < 15 ? 0x100u : 0u);
        present[w] = (uint16_t)p;
        if (pf | pr) atomicOr(bloom, pf | pr);
        if (len < 15) atomicOr(bloom, 0x100u);
    }
}
/* map_to_index (:388-529) per orientation: votes exist <=> the read holds a key with 1..50 positions of which one packs to a
 * non-zero value (`voting`); then the mask loop panics <=> the read also holds a valid base whose key is absent (`absent`).
 * first[0] = smallest (seq index * 2 + orientation) that panics. */
__global__ void k_seq_decide(const uint16_t* __restrict__ present, uint64_t n, uint32_t voting, uint32_t absent,
                             unsigned long long* __restrict__ first) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t p = present[j], f = p & 15u, r = (p >> 4) & 15u;
    if ((f & voting) && (f & absent)) atomicMin(first, 2ull * j);
    else if ((r & voting) && (r & absent)) atomicMin(first, 2ull * j + 1ull);
}

struct PinnedBuf {
    void* p = nullptr;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
struct EvGuard {
    cudaEvent_t e = nullptr;
    ~EvGuard() { if (e) cudaEventDestroy(e); }
};

}  // namespace

struct gf_reference {
    int device = 0;
    gf_reference_info info{};
    std::vector<std::pair<uint32_t, uint32_t>> listed[4]; /* (contig, position) ascending, keys with <= RS_LIST_CAP positions */
    std::mutex mu;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
};

namespace {

/* one launch over a set of segments; the table is uploaded on `st` from pinned staging owned by the caller */
template <bool EMIT>
void launch_scan(const RefSeg* d_segs, uint32_t n_segs, uint64_t total_wt, RefScanOut* d_out, uint32_t emit_mask, int sms,
                 cudaStream_t st) {
    if (!total_wt) return;
    const uint64_t blocks_needed = (total_wt + RS_THREADS / 32 - 1) / (RS_THREADS / 32);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ref_scan<EMIT>, RS_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    const unsigned grid = (unsigned)std::min<uint64_t>(blocks_needed, (uint64_t)sms * per_sm);
    k_ref_scan<EMIT><<<grid, RS_THREADS, 0, st>>>(d_segs, n_segs, total_wt, d_out, emit_mask);
}

/* warp tiles of a segment: its 32-byte sectors in groups of 31 */
uint64_t seg_warp_tiles(const uint8_t* p, uint64_t n) {
    if (!n) return 0;
    const uintptr_t first = (uintptr_t)p & ~(uintptr_t)31, last = ((uintptr_t)p + n + 31) & ~(uintptr_t)31;
    const uint64_t sectors = (last - first) / 32;
    return (sectors + RS_WT_SECTORS - 1) / RS_WT_SECTORS;
}

/* the whole pass: emit_mask == 0 counts, otherwise lists the positions of the keys in the mask */
int scan_reference(gf_reference* ref, const gf_ref_contig* contigs, uint32_t n_contigs, uint32_t emit_mask, RefScanOut* h_out,
                   float* ms_scan, uint64_t* h2d_bytes, uint64_t* launches) {
    cudaStream_t st = ref->stream;
    DevBuf d_out, d_stage[2], d_segs[2];
    PinnedBuf h_segs[2], h_res;
    EvGuard ev_free[2], ev_k0, ev_k1;
    GF_CUDA_TRY(cudaMalloc(&d_out.p, sizeof(RefScanOut)));
    GF_CUDA_TRY(cudaMemsetAsync(d_out.p, 0, sizeof(RefScanOut), st));
    GF_CUDA_TRY(cudaMallocHost(&h_res.p, sizeof(RefScanOut)));
    GF_CUDA_TRY(cudaEventCreate(&ev_k0.e));
    GF_CUDA_TRY(cudaEventCreate(&ev_k1.e));
    const size_t max_segs = 4096; /* a buffer is flushed when its segment table is full */
    bool any_host = false;
    for (uint32_t c = 0; c < n_contigs; c++) {
        if (!contigs[c].len) continue;
        cudaPointerAttributes at;
        bool dev = cudaPointerGetAttributes(&at, contigs[c].seq) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        if (!dev) any_host = true;
    }
    for (int k = 0; k < 2; k++) {
        if (any_host) GF_CUDA_TRY(cudaMalloc(&d_stage[k].p, RS_CHUNK + 256));
        GF_CUDA_TRY(cudaMalloc(&d_segs[k].p, sizeof(RefSeg) * max_segs));
        GF_CUDA_TRY(cudaMallocHost(&h_segs[k].p, sizeof(RefSeg) * max_segs));
        GF_CUDA_TRY(cudaEventCreate(&ev_free[k].e));
    }
    float ms_k = 0;
    int cur = 0;
    bool used[2] = {false, false};
    std::vector<RefSeg> segs;
    uint64_t total_wt = 0;
    uint64_t fill = 0; /* bytes used in the current staging buffer */
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed; /* kernel start/end events, read at the end */
    std::vector<EvGuard> ev_pool;
    ev_pool.reserve(4096);

    auto flush = [&]() -> int {
        if (segs.empty()) return GF_OK;
        memcpy(h_segs[cur].p, segs.data(), sizeof(RefSeg) * segs.size());
        GF_CUDA_TRY(cudaMemcpyAsync(d_segs[cur].p, h_segs[cur].p, sizeof(RefSeg) * segs.size(), cudaMemcpyHostToDevice, st));
        cudaEvent_t a = nullptr, b = nullptr;
        if (ev_pool.size() + 2 <= ev_pool.capacity()) {
            ev_pool.emplace_back(); GF_CUDA_TRY(cudaEventCreate(&ev_pool.back().e)); a = ev_pool.back().e;
            ev_pool.emplace_back(); GF_CUDA_TRY(cudaEventCreate(&ev_pool.back().e)); b = ev_pool.back().e;
            GF_CUDA_TRY(cudaEventRecord(a, st));
        }
        if (emit_mask) launch_scan<true>((const RefSeg*)d_segs[cur].p, (uint32_t)segs.size(), total_wt, (RefScanOut*)d_out.p, emit_mask,
                                         ref->sm_count, st);
        else launch_scan<false>((const RefSeg*)d_segs[cur].p, (uint32_t)segs.size(), total_wt, (RefScanOut*)d_out.p, 0u, ref->sm_count, st);
        GF_CUDA_TRY(cudaGetLastError());
        if (b) { GF_CUDA_TRY(cudaEventRecord(b, st)); timed.emplace_back(a, b); }
        (*launches)++;
        GF_CUDA_TRY(cudaEventRecord(ev_free[cur].e, st));
        used[cur] = true;
        segs.clear();
        total_wt = 0;
        fill = 0;
        cur ^= 1;
        /* the buffer we switch to (staging + pinned tables) may still be in use by the launch before last */
        if (used[cur]) GF_CUDA_TRY(cudaEventSynchronize(ev_free[cur].e));
        return GF_OK;
    };

    for (uint32_t c = 0; c < n_contigs; c++) {
        const uint64_t len = contigs[c].len;
        if (len == 0) continue;
        cudaPointerAttributes at;
        const bool dev = cudaPointerGetAttributes(&at, contigs[c].seq) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        if (dev) { /* resident: scanned in place, one segment */
            if (segs.size() + 1 > max_segs) { int r = flush(); if (r) return r; }
            const uint64_t nwt = seg_warp_tiles(contigs[c].seq, len);
            segs.push_back(RefSeg{contigs[c].seq, len, 0, len, 0, c, total_wt, nwt});
            total_wt += nwt;
            continue;
        }
        uint64_t done = 0;
        while (done < len) {
            const uint32_t lead = (uint32_t)std::min<uint64_t>(done, 16);
            /* staged layout: [.. fill) used; this piece goes to a 32-byte aligned start + 32 so that `lead` bytes fit before */
            uint64_t start = ((fill + 31) & ~31ull) + 32;
            if (start + 64 > RS_CHUNK || segs.size() + 1 > max_segs) {
                int r = flush();
                if (r) return r;
                start = 32;
            }
            const uint64_t n = std::min<uint64_t>(len - done, RS_CHUNK - start);
            uint8_t* dst = (uint8_t*)d_stage[cur].p + start;
            GF_CUDA_TRY(cudaMemcpyAsync(dst - lead, contigs[c].seq + done - lead, n + lead, cudaMemcpyHostToDevice, st));
            *h2d_bytes += n + lead;
            const uint64_t nwt = seg_warp_tiles(dst, n);
            segs.push_back(RefSeg{dst, n, done, len, lead, c, total_wt, nwt});
            total_wt += nwt;
            fill = start + n;
            done += n;
        }
    }
    { int r = flush(); if (r) return r; }
    GF_CUDA_TRY(cudaMemcpyAsync(h_res.p, d_out.p, sizeof(RefScanOut), cudaMemcpyDeviceToHost, st));
    GF_CUDA_TRY(cudaStreamSynchronize(st));
    for (auto& ab : timed) {
        float ms = 0;
        GF_CUDA_TRY(cudaEventElapsedTime(&ms, ab.first, ab.second));
        ms_k += ms;
    }
    *ms_scan += ms_k;
    memcpy(h_out, h_res.p, sizeof(RefScanOut));
    return GF_OK;
}

}  // namespace

extern "C" {

int gf_reference_create(const gf_ref_contig* contigs, uint32_t n_contigs, int device, gf_reference** out) {
    if (!out) { gf_set_error("out is NULL"); return GF_E_INVALID; }
    *out = nullptr;
    if (n_contigs && !contigs) { gf_set_error("contigs is NULL"); return GF_E_INVALID; }
    if (n_contigs > 32767) { gf_set_error("more than 32767 contigs: contig ids are i16 (src/core/common.rs:5)"); return GF_E_LIMIT; }
    for (uint32_t c = 0; c < n_contigs; c++) {
        if (contigs[c].len && !contigs[c].seq) { gf_set_error("contig with len > 0 and seq == NULL"); return GF_E_INVALID; }
        if (contigs[c].len > 0x7FFFFFFFull) { gf_set_error("contig longer than 2^31-1 bases: positions are i32"); return GF_E_LIMIT; }
    }
    if (gf_device_count() <= 0) { gf_set_error("no CUDA device available (this library has no CPU fallback)"); return GF_E_CUDA; }
    if (device < 0 || device >= gf_device_count()) { gf_set_error("device index out of range"); return GF_E_INVALID; }
    GF_CUDA_TRY(cudaSetDevice(device));
    gf_reference* ref = new gf_reference();
    ref->device = device;
    struct Cleanup { gf_reference* r; ~Cleanup() { if (r) { if (r->stream) cudaStreamDestroy(r->stream); delete r; } } } guard{ref};
    GF_CUDA_TRY(cudaDeviceGetAttribute(&ref->sm_count, cudaDevAttrMultiProcessorCount, device));
    GF_CUDA_TRY(cudaStreamCreateWithFlags(&ref->stream, cudaStreamNonBlocking));
    EvGuard e0, e1;
    GF_CUDA_TRY(cudaEventCreate(&e0.e));
    GF_CUDA_TRY(cudaEventCreate(&e1.e));
    gf_reference_info& inf = ref->info;
    inf.n_contigs = n_contigs;
    for (uint32_t c = 0; c < n_contigs; c++) {
        inf.n_bases += contigs[c].len;
        if (contigs[c].len < 16) inf.short_contigs++;
    }
    GF_CUDA_TRY(cudaEventRecord(e0.e, ref->stream));
    RefScanOut res;
    int rc = scan_reference(ref, contigs, n_contigs, 0u, &res, &inf.ms_scan, &inf.h2d_bytes, &inf.kernel_launches);
    if (rc != GF_OK) return rc;
    uint32_t emit = 0;
    for (int k = 0; k < 4; k++) {
        inf.key_positions[k] = res.count[k];
        if (res.count[k] >= 1 && res.count[k] <= (unsigned long long)RS_LIST_CAP) emit |= 1u << k;
    }
    if (emit) { /* a key with few positions: map_to_index votes with the positions themselves */
        RefScanOut lst;
        uint64_t h2d2 = 0;
        rc = scan_reference(ref, contigs, n_contigs, emit, &lst, &inf.ms_scan, &h2d2, &inf.kernel_launches);
        if (rc != GF_OK) return rc;
        inf.h2d_bytes += h2d2;
        for (int k = 0; k < 4; k++) {
            if (!((emit >> k) & 1u)) continue;
            if (lst.n_listed[k] != res.count[k]) { gf_set_error("internal: reference scan passes disagree"); return GF_E_CUDA; }
            for (unsigned j = 0; j < lst.n_listed[k]; j++) ref->listed[k].emplace_back(lst.list_contig[k][j], lst.list_pos[k][j]);
            std::sort(ref->listed[k].begin(), ref->listed[k].end()); /* single-thread push order: contig, then position */
        }
    }
    GF_CUDA_TRY(cudaEventRecord(e1.e, ref->stream));
    GF_CUDA_TRY(cudaEventSynchronize(e1.e));
    GF_CUDA_TRY(cudaEventElapsedTime(&inf.ms_total, e0.e, e1.e));
    guard.r = nullptr;
    *out = ref;
    return GF_OK;
}

void gf_reference_destroy(gf_reference* ref) {
    if (!ref) return;
    cudaSetDevice(ref->device);
    if (ref->stream) { cudaStreamSynchronize(ref->stream); cudaStreamDestroy(ref->stream); }
    delete ref;
}

int gf_reference_get_info(const gf_reference* ref, gf_reference_info* out) {
    if (!ref || !out) { gf_set_error("NULL argument"); return GF_E_INVALID; }
    *out = ref->info;
    return GF_OK;
}

int gf_alignable_filter(gf_reference* ref, const uint8_t* seqs, const uint64_t* seq_off, uint64_t n_seqs, uint8_t* alignable,
                        gf_alignable_result* res) {
    if (!ref || !res || (n_seqs && (!seq_off || !alignable))) { gf_set_error("NULL argument"); return GF_E_INVALID; }
    memset(res, 0, sizeof(*res));
    res->panic_seq = -1;
    if (n_seqs) memset(alignable, 0, n_seqs);
    const uint64_t n_bytes = n_seqs ? seq_off[n_seqs] - seq_off[0] : 0;
    if (n_bytes && !seqs) { gf_set_error("seqs is NULL"); return GF_E_INVALID; }
    for (uint64_t j = 0; j < n_seqs; j++)
        if (seq_off[j + 1] < seq_off[j]) { gf_set_error("offsets are not ascending"); return GF_E_INVALID; }
    std::lock_guard<std::mutex> lk(ref->mu);
    GF_CUDA_TRY(cudaSetDevice(ref->device));
    cudaStream_t st = ref->stream;
    uint32_t bloom = 0;
    DevBuf d_seqs, d_off, d_present, d_small;
    if (n_seqs) {
        GF_CUDA_TRY(cudaMalloc(&d_seqs.p, n_bytes + 16));
        GF_CUDA_TRY(cudaMalloc(&d_off.p, sizeof(uint64_t) * (n_seqs + 1)));
        GF_CUDA_TRY(cudaMalloc(&d_present.p, sizeof(uint16_t) * n_seqs));
        GF_CUDA_TRY(cudaMalloc(&d_small.p, 64));
        GF_CUDA_TRY(cudaMemsetAsync(d_small.p, 0, 64, st));
        if (n_bytes) GF_CUDA_TRY(cudaMemcpyAsync(d_seqs.p, seqs + seq_off[0], n_bytes, cudaMemcpyHostToDevice, st));
        /* offsets relative to the first byte */
        std::vector<uint64_t> rel(n_seqs + 1);
        for (uint64_t j = 0; j <= n_seqs; j++) rel[j] = seq_off[j] - seq_off[0];
        GF_CUDA_TRY(cudaMemcpyAsync(d_off.p, rel.data(), sizeof(uint64_t) * (n_seqs + 1), cudaMemcpyHostToDevice, st));
        GF_CUDA_TRY(cudaStreamSynchronize(st)); /* `rel` is pageable host memory */
        const unsigned blocks = (unsigned)((n_seqs * 32 + 255) / 256);
        k_seq_present<<<blocks, 256, 0, st>>>((const uint8_t*)d_seqs.p, (const unsigned long long*)d_off.p, n_seqs,
                                               (uint16_t*)d_present.p, (unsigned int*)d_small.p);
        GF_CUDA_TRY(cudaGetLastError());
        GF_CUDA_TRY(cudaMemcpyAsync(&bloom, d_small.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        GF_CUDA_TRY(cudaStreamSynchronize(st));
    }
    res->bloom_bits = bloom & 15u;
    uint32_t voting = 0, absent = 0;
    for (int k = 0; k < 4; k++) {
        const uint64_t nk = ((bloom >> k) & 1u) ? ref->info.key_positions[k] : 0;
        res->key_positions[k] = nk;
        if (nk == 0) { absent |= 1u << k; continue; }
        if (nk > 50) continue; /* skip_threshold (:404, :426-429): no votes */
        /* votes are added under pack(contig, position - j), j = index in the key's list (:432-433); the value 0 is
         * excluded from the top-5 (:450), i.e. an entry counts unless contig == 0 and position == j */
        const auto& L = ref->listed[k];
        for (size_t j = 0; j < L.size(); j++)
            if (!(L[j].first == 0 && L[j].second == (uint32_t)j)) { voting |= 1u << k; break; }
    }
    if (bloom & 0x100u) { /* a sequence shorter than 15 bases: init_bloom_filter's range arithmetic wraps (:77) */
        res->panic_stage = 4;
        for (uint64_t j = 0; j < n_seqs; j++)
            if (seq_off[j + 1] - seq_off[j] < 15) { res->panic_seq = (int64_t)j; break; }
        gf_set_error("a sequence is shorter than 15 bases: the reference's Matcher is undefined there (matcher.rs:77)");
        return GF_E_REF_PANIC;
    }
    if (ref->info.short_contigs) {
        res->panic_stage = 1;
        gf_set_error("a reference contig is shorter than 16 bases: Matcher::make_index panics (matcher.rs:240-243)");
        return GF_E_REF_PANIC;
    }
    if (n_seqs && voting && absent) {
        unsigned long long first = ~0ull;
        unsigned long long* d_first = (unsigned long long*)((uint8_t*)d_small.p + 8);
        GF_CUDA_TRY(cudaMemsetAsync(d_first, 0xFF, sizeof(unsigned long long), st));
        k_seq_decide<<<(unsigned)((n_seqs + 255) / 256), 256, 0, st>>>((const uint16_t*)d_present.p, n_seqs, voting, absent, d_first);
        GF_CUDA_TRY(cudaGetLastError());
        GF_CUDA_TRY(cudaMemcpyAsync(&first, d_first, sizeof(first), cudaMemcpyDeviceToHost, st));
        GF_CUDA_TRY(cudaStreamSynchronize(st));
        if (first != ~0ull) {
            res->panic_seq = (int64_t)(first >> 1);
            res->panic_stage = (first & 1ull) ? 3 : 2;
            gf_set_error("Matcher::map_to_index unwraps a missing key for this sequence: the reference panics (matcher.rs:490-491)");
            return GF_E_REF_PANIC;
        }
    }
    return GF_OK;
}

} /* extern "C" */
