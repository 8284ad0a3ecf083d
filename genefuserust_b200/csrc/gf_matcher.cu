/*
 * gf_matcher.cu — the Matcher pass of FusionMapper::remove_alignables (/root/reference/src/core/fusion_mapper.rs:488-542,
 * src/core/matcher.rs) as ONE streaming scan of the reference on the GPU.  include/genefuse_gpu.h says what the (degenerate)
 * reference code computes; here is how it maps to kernels.
 *
 *   k_ref_scan     thread per aligned 32-byte sector of reference text: 32 ASCII bases -> three plane words (code lo / hi,
 *                  valid; case-insensitive = to_ascii_uppercase, matcher.rs:143-148) with the SWAR converters of gf_swar.cuh,
 *                  neighbour words exchanged through shared memory, then the keep rule of index_contig_bytes (:227-289) for
 *                  all 32 positions at once in the bit domain:
 *                      position i is kept  <=>  base i is ACGT, i < len - 16, and walking back from i-1 over at most 15
 *                      positions the first base that is not 'A' is a non-ACGT byte (or lies before the contig start), or
 *                      all 15 are 'A'                      (<=> the rolling 32-bit value of the run is < 4)
 *                  and its key is the base's own 2-bit code.  Per block: four counters, one atomicAdd each.
 *                  HBM streaming: 1 byte per base read, nothing written.
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

constexpr int RS_THREADS = 128;            /* sectors (32 bases) per tile */
constexpr uint64_t RS_CHUNK = 64ull << 20; /* staging buffer bytes (x2, double buffered) */
constexpr int RS_LIST_CAP = 64;            /* positions kept per key (votes need at most 50) */

struct RefSeg {      /* a run of bases of one contig, resident at device address p */
    const uint8_t* p;    /* first COUNTED base */
    uint64_t n;          /* counted bases */
    uint64_t pos0;       /* contig position of p[0] */
    uint64_t contig_len;
    uint32_t lead;       /* readable context bytes before p (>= 16 unless pos0 == 0) */
    uint32_t contig;
};
struct RefTile {
    uint64_t sector;     /* address of the tile's first 32-byte sector */
    uint32_t seg;
    uint32_t pad;
};
struct RefScanOut {
    unsigned long long count[4];
    unsigned int n_listed[4];
    unsigned int list_contig[4][RS_LIST_CAP];
    unsigned int list_pos[4][RS_LIST_CAP];
};

__device__ __forceinline__ void load_sector(const uint8_t* sec, const uint8_t* lo, const uint8_t* hi, uint4* a, uint4* b) {
    if (sec >= lo && sec + 32 <= hi) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w)
                     : "l"(sec));
        return;
    }
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sec + 32 > lo && sec < hi)
        for (int t = 0; t < 32; t++)
            if (sec + t >= lo && sec + t < hi) w[t >> 2] |= (uint32_t)__ldg(sec + t) << (8 * (t & 3));
    *a = make_uint4(w[0], w[1], w[2], w[3]);
    *b = make_uint4(w[4], w[5], w[6], w[7]);
}
/* 32 bases -> plane words (bit t = byte t of the sector); bytes outside the segment were read as 0 = invalid */
__device__ __forceinline__ void sector_planes(const uint4& a, const uint4& b, uint32_t* lo, uint32_t* hi, uint32_t* v) {
    uint32_t l0, h0, v0, e0, l1, h1, v1, e1;
    swar::block16<true>(a, &l0, &h0, &v0, &e0);
    swar::block16<true>(b, &l1, &h1, &v1, &e1);
    *lo = __byte_perm(l0, l1, 0x5410u);
    *hi = __byte_perm(h0, h1, 0x5410u);
    *v = __byte_perm(v0, v1, 0x5410u);
}

template <bool EMIT>
__global__ void __launch_bounds__(RS_THREADS) k_ref_scan(const RefSeg* __restrict__ segs, const RefTile* __restrict__ tiles,
                                                         uint32_t n_tiles, RefScanOut* __restrict__ out, uint32_t emit_mask) {
    __shared__ uint32_t s_lo[RS_THREADS + 1], s_hi[RS_THREADS + 1], s_v[RS_THREADS + 1];
    __shared__ unsigned int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (uint32_t ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
        const RefTile tile = tiles[ti];
        const RefSeg sg = segs[tile.seg];
        const uint8_t* lo_b = sg.p - sg.lead;
        const uint8_t* hi_b = sg.p + sg.n;
        const uint8_t* sec = reinterpret_cast<const uint8_t*>(tile.sector) + 32ull * threadIdx.x;
        uint4 a, b;
        uint32_t lo, hi, v;
        load_sector(sec, lo_b, hi_b, &a, &b);
        sector_planes(a, b, &lo, &hi, &v);
        __syncthreads(); /* the previous tile's readers are done */
        s_lo[threadIdx.x + 1] = lo; s_hi[threadIdx.x + 1] = hi; s_v[threadIdx.x + 1] = v;
        if (threadIdx.x == 0) { /* the sector before the tile: context for its first 15 positions */
            uint4 pa, pb;
            uint32_t plo, phi, pv;
            load_sector(sec - 32, lo_b, hi_b, &pa, &pb);
            sector_planes(pa, pb, &plo, &phi, &pv);
            s_lo[0] = plo; s_hi[0] = phi; s_v[0] = pv;
        }
        __syncthreads();
        const unsigned long long v64 = ((unsigned long long)v << 32) | s_v[threadIdx.x];
        const unsigned long long lo64 = ((unsigned long long)lo << 32) | s_lo[threadIdx.x];
        const unsigned long long hi64 = ((unsigned long long)hi << 32) | s_hi[threadIdx.x];
        /* blocked[i] <=> some valid non-'A' base (a generate G) lies d <= 15 positions behind i with nothing but 'A's (propagate
         * P) in between: a carry look-ahead over a window of 15, in 3 doubling steps + 3 combines (bit b of X << d = position
         * i - d).  Non-ACGT bytes neither generate nor propagate, so a run start is never blocked; 15 'A's are a window without
         * a generate.  kept = valid & ~blocked. */
        const unsigned long long P1 = v64 & ~lo64 & ~hi64, G1 = v64 & ~P1;
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
        const uint32_t kept = (uint32_t)((v64 & ok) >> 32) & cm;
        const uint32_t kA = kept & ~lo & ~hi, kT = kept & lo & ~hi, kC = kept & ~lo & hi, kG = kept & lo & hi;
        c0 += __popc(kA); c1 += __popc(kT); c2 += __popc(kC); c3 += __popc(kG);
        if (EMIT) {
            const uint32_t km[4] = {kA, kT, kC, kG};
            for (int k = 0; k < 4; k++) {
                if (!((emit_mask >> k) & 1u)) continue;
                uint32_t m = km[k];
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned slot = atomicAdd(&out->n_listed[k], 1u);
                    if (slot < (unsigned)RS_LIST_CAP) {
                        out->list_contig[k][slot] = sg.contig;
                        out->list_pos[k][slot] = (unsigned int)(sg.pos0 + (uint64_t)(first + bit));
                    }
                }
            }
        }
    }
    c0 = __reduce_add_sync(0xFFFFFFFFu, c0); c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
    c2 = __reduce_add_sync(0xFFFFFFFFu, c2); c3 = __reduce_add_sync(0xFFFFFFFFu, c3);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
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

/* one launch over a set of tiles; the tables are uploaded on `st` from pinned staging owned by the caller */
template <bool EMIT>
void launch_scan(const RefSeg* d_segs, const RefTile* d_tiles, uint32_t n_tiles, RefScanOut* d_out, uint32_t emit_mask, int sms,
                 cudaStream_t st) {
    if (!n_tiles) return;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sms * 16);
    k_ref_scan<EMIT><<<grid, RS_THREADS, 0, st>>>(d_segs, d_tiles, n_tiles, d_out, emit_mask);
}

void add_tiles(std::vector<RefTile>& tiles, uint32_t seg_id, const uint8_t* p, uint64_t n, uint32_t lead) {
    (void)lead;
    if (!n) return;
    const uintptr_t first = (uintptr_t)p & ~(uintptr_t)31, last = ((uintptr_t)p + n + 31) & ~(uintptr_t)31;
    for (uintptr_t s = first; s < last; s += 32ull * RS_THREADS) tiles.push_back(RefTile{(uint64_t)s, seg_id, 0});
}

/* the whole pass: emit_mask == 0 counts, otherwise lists the positions of the keys in the mask */
int scan_reference(gf_reference* ref, const gf_ref_contig* contigs, uint32_t n_contigs, uint32_t emit_mask, RefScanOut* h_out,
                   float* ms_scan, uint64_t* h2d_bytes, uint64_t* launches) {
    cudaStream_t st = ref->stream;
    DevBuf d_out, d_stage[2], d_segs[2], d_tiles[2];
    PinnedBuf h_segs[2], h_tiles[2], h_res;
    EvGuard ev_free[2], ev_k0, ev_k1;
    GF_CUDA_TRY(cudaMalloc(&d_out.p, sizeof(RefScanOut)));
    GF_CUDA_TRY(cudaMemsetAsync(d_out.p, 0, sizeof(RefScanOut), st));
    GF_CUDA_TRY(cudaMallocHost(&h_res.p, sizeof(RefScanOut)));
    GF_CUDA_TRY(cudaEventCreate(&ev_k0.e));
    GF_CUDA_TRY(cudaEventCreate(&ev_k1.e));
    /* upper bounds of the per-buffer tables: a tile covers 4096 bytes; segments are >= 1 byte but a buffer holds at most
     * max_segs of them (the buffer is flushed when the table is full) */
    const size_t max_segs = 4096, max_tiles = (size_t)(RS_CHUNK / (32 * RS_THREADS)) + 2 * max_segs + 16;
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
        GF_CUDA_TRY(cudaMalloc(&d_tiles[k].p, sizeof(RefTile) * max_tiles));
        GF_CUDA_TRY(cudaMallocHost(&h_segs[k].p, sizeof(RefSeg) * max_segs));
        GF_CUDA_TRY(cudaMallocHost(&h_tiles[k].p, sizeof(RefTile) * max_tiles));
        GF_CUDA_TRY(cudaEventCreate(&ev_free[k].e));
    }
    float ms_k = 0;
    int cur = 0;
    bool used[2] = {false, false};
    std::vector<RefSeg> segs;
    std::vector<RefTile> tiles;
    uint64_t fill = 0; /* bytes used in the current staging buffer */
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed; /* kernel start/end events, read at the end */
    std::vector<EvGuard> ev_pool;
    ev_pool.reserve(4096);

    auto flush = [&]() -> int {
        if (segs.empty()) return GF_OK;
        memcpy(h_segs[cur].p, segs.data(), sizeof(RefSeg) * segs.size());
        memcpy(h_tiles[cur].p, tiles.data(), sizeof(RefTile) * tiles.size());
        GF_CUDA_TRY(cudaMemcpyAsync(d_segs[cur].p, h_segs[cur].p, sizeof(RefSeg) * segs.size(), cudaMemcpyHostToDevice, st));
        GF_CUDA_TRY(cudaMemcpyAsync(d_tiles[cur].p, h_tiles[cur].p, sizeof(RefTile) * tiles.size(), cudaMemcpyHostToDevice, st));
        cudaEvent_t a = nullptr, b = nullptr;
        if (ev_pool.size() + 2 <= ev_pool.capacity()) {
            ev_pool.emplace_back(); GF_CUDA_TRY(cudaEventCreate(&ev_pool.back().e)); a = ev_pool.back().e;
            ev_pool.emplace_back(); GF_CUDA_TRY(cudaEventCreate(&ev_pool.back().e)); b = ev_pool.back().e;
            GF_CUDA_TRY(cudaEventRecord(a, st));
        }
        if (emit_mask) launch_scan<true>((const RefSeg*)d_segs[cur].p, (const RefTile*)d_tiles[cur].p, (uint32_t)tiles.size(),
                                         (RefScanOut*)d_out.p, emit_mask, ref->sm_count, st);
        else launch_scan<false>((const RefSeg*)d_segs[cur].p, (const RefTile*)d_tiles[cur].p, (uint32_t)tiles.size(),
                                (RefScanOut*)d_out.p, 0u, ref->sm_count, st);
        GF_CUDA_TRY(cudaGetLastError());
        if (b) { GF_CUDA_TRY(cudaEventRecord(b, st)); timed.emplace_back(a, b); }
        (*launches)++;
        GF_CUDA_TRY(cudaEventRecord(ev_free[cur].e, st));
        used[cur] = true;
        segs.clear();
        tiles.clear();
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
            if (segs.size() + 1 > max_segs || tiles.size() + len / (32 * RS_THREADS) + 2 > max_tiles) { int r = flush(); if (r) return r; }
            uint64_t done = 0;
            while (done < len) { /* cut only so that the tile table fits */
                const uint64_t room = (max_tiles - tiles.size() - 2) * 32ull * RS_THREADS;
                const uint64_t n = std::min<uint64_t>(len - done, room);
                segs.push_back(RefSeg{contigs[c].seq + done, n, done, len, (uint32_t)std::min<uint64_t>(done, 16), c});
                add_tiles(tiles, (uint32_t)segs.size() - 1, contigs[c].seq + done, n, 0);
                done += n;
                if (done < len) { int r = flush(); if (r) return r; }
            }
            continue;
        }
        uint64_t done = 0;
        while (done < len) {
            const uint32_t lead = (uint32_t)std::min<uint64_t>(done, 16);
            /* staged layout: [.. fill) used; this piece goes to a 32-byte aligned start + 32 so that `lead` bytes fit before */
            uint64_t start = ((fill + 31) & ~31ull) + 32;
            if (start + 64 > RS_CHUNK || segs.size() + 1 > max_segs || tiles.size() + 4 > max_tiles) {
                int r = flush();
                if (r) return r;
                start = 32;
            }
            const uint64_t n = std::min<uint64_t>(len - done, RS_CHUNK - start);
            uint8_t* dst = (uint8_t*)d_stage[cur].p + start;
            GF_CUDA_TRY(cudaMemcpyAsync(dst - lead, contigs[c].seq + done - lead, n + lead, cudaMemcpyHostToDevice, st));
            *h2d_bytes += n + lead;
            segs.push_back(RefSeg{dst, n, done, len, lead, c});
            add_tiles(tiles, (uint32_t)segs.size() - 1, dst, n, lead);
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
