/*
 * gf_fastq.cu — FASTQ ingest on the device: raw FASTQ text -> per-record tables, so the
 * mapping kernels read sequences and qualities straight out of the text buffer (no host-side parsing, no repacking).
 *
 * Mirrors FastqReader::read (/root/reference/src/core/fastq_reader.rs:75-147): a record is four lines (name,
 * sequence, strand, quality), each line loses ONE trailing '\n' and nothing else ('\r' is kept, nothing is validated);
 * reading stops at the first record that cannot be completed, so an incomplete trailing record is dropped.
 *
 *   k_nl_count   16 bytes per thread (128-bit loads), newlines per 4 KB tile
 *   exclusive scan of the tile counts (gf_exclusive_scan_u32)
 *   k_nl_write   positions of all newlines, in order
 *   k_records    record i: sequence = (NL[4i], NL[4i+1]), quality = (NL[4i+2], NL[4i+3]); max length; checks
 */
#include "gf_internal.h"

namespace {

constexpr int NL_THREADS = 256;
constexpr int NL_TILE = NL_THREADS * 16;

__device__ __forceinline__ uint32_t newline_mask16(const uint8_t* __restrict__ text, uint64_t bytes, uint64_t pos) {
    /* bit t set <=> text[pos + t] == '\n' */
    uint32_t m = 0;
    if (pos + 16 <= bytes && ((uintptr_t)(text + pos) & 15u) == 0) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(text + pos));
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t y = w[j] ^ 0x0A0A0A0Au;
            uint32_t z = ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u; /* bit 7 of every zero byte */
            m |= ((((z >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * j);
        }
    } else {
        for (int t = 0; t < 16; t++)
            if (pos + t < bytes && text[pos + t] == '\n') m |= 1u << t;
    }
    return m;
}
__global__ void k_nl_count(const uint8_t* __restrict__ text, uint64_t bytes, uint32_t* __restrict__ tile_counts,
                           unsigned long long* __restrict__ total) {
    __shared__ uint32_t ws[NL_THREADS / 32];
    uint64_t pos = (uint64_t)blockIdx.x * NL_TILE + (uint64_t)threadIdx.x * 16;
    uint32_t c = pos < bytes ? __popc(newline_mask16(text, bytes, pos)) : 0u;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < NL_THREADS / 32; w++) t += ws[w];
        tile_counts[blockIdx.x] = t;
        if (t) atomicAdd(total, (unsigned long long)t); /* 64-bit: the u32 scan below must not wrap unnoticed */
    }
}
__global__ void k_nl_write(const uint8_t* __restrict__ text, uint64_t bytes, const uint32_t* __restrict__ tile_off,
                           unsigned long long* __restrict__ nl) {
    __shared__ uint32_t ws[NL_THREADS / 32];
    uint64_t pos = (uint64_t)blockIdx.x * NL_TILE + (uint64_t)threadIdx.x * 16;
    uint32_t m = pos < bytes ? newline_mask16(text, bytes, pos) : 0u;
    uint32_t c = __popc(m), incl = c;
    const uint32_t lane = threadIdx.x & 31;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += ws[w];
    uint32_t dst = tile_off[blockIdx.x] + woff + incl - c;
    while (m) {
        int t = __ffs(m) - 1;
        m &= m - 1;
        nl[dst++] = pos + (uint64_t)t;
    }
}
/* err bits: 1 = quality and sequence lengths differ */
__global__ void k_records(const unsigned long long* __restrict__ nl, uint64_t n_records, unsigned long long* __restrict__ s,
                          unsigned long long* __restrict__ e, unsigned long long* __restrict__ qs,
                          unsigned int* __restrict__ max_len, unsigned int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0, bad = 0;
    if (i < n_records) {
        /* line j spans (nl[j], nl[j+1]) in the shifted table; record i = lines 4i (name), 4i+1 (sequence), 4i+2, 4i+3 (quality) */
        unsigned long long a = nl[4 * i + 1] + 1, b = nl[4 * i + 2], qa = nl[4 * i + 3] + 1, qb = nl[4 * i + 4];
        s[i] = a;
        e[i] = b;
        qs[i] = qa;
        len = (uint32_t)(b - a);
        bad = (qb - qa) != (b - a);
    }
    len = __reduce_max_sync(0xFFFFFFFFu, len);
    bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (len) atomicMax(max_len, len);
        if (bad) atomicOr(err, 1u);
    }
}

}  // namespace

int gf_fastq_parse_device(const uint8_t* d_text, uint64_t bytes, GfFastqTable* out, cudaStream_t st, bool final_chunk) {
    out->n_records = 0;
    out->max_len = 0;
    if (bytes == 0) return GF_OK;
    const uint64_t n_tiles = (bytes + NL_TILE - 1) / NL_TILE;
    if (n_tiles > 0x7FFFFFFFull) { gf_set_error("FASTQ buffer too large for one call (> 8 TiB)"); return GF_E_LIMIT; }
    GF_CUDA_TRY(out->cnt.reserve(sizeof(uint32_t) * (n_tiles + 2) + 16));
    GF_CUDA_TRY(out->off.reserve(sizeof(uint32_t) * (n_tiles + 1)));
    GF_CUDA_TRY(out->tmp.reserve(sizeof(uint32_t) * gf_scan_tmp_elems(n_tiles)));
    uint32_t *d_cnt = out->cnt.as<uint32_t>(), *d_off = out->off.as<uint32_t>(), *d_tmp = out->tmp.as<uint32_t>();
    /* 64-bit newline total behind the tile counts (8-byte aligned) */
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_cnt + ((n_tiles + 2 + 1) & ~1ull));
    GF_CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
    k_nl_count<<<(unsigned)n_tiles, NL_THREADS, 0, st>>>(d_text, bytes, d_cnt, d_total);
    GF_CUDA_TRY(gf_exclusive_scan_u32(d_cnt, d_off, n_tiles, d_tmp, st));
    uint32_t n_nl = 0;
    unsigned long long n_nl64 = 0;
    uint8_t last = 0;
    GF_CUDA_TRY(cudaMemcpyAsync(&n_nl, d_off + n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GF_CUDA_TRY(cudaMemcpyAsync(&n_nl64, d_total, sizeof(n_nl64), cudaMemcpyDeviceToHost, st));
    GF_CUDA_TRY(cudaMemcpyAsync(&last, d_text + bytes - 1, 1, cudaMemcpyDeviceToHost, st));
    GF_CUDA_TRY(cudaStreamSynchronize(st));
    if (n_nl64 != (unsigned long long)n_nl) {
        gf_set_error("FASTQ buffer holds 2^32 or more lines: pass it in pieces (gf_fastq_stream_*)");
        return GF_E_LIMIT;
    }
    /* a non-empty unterminated last line still counts as a line (read_line returns > 0): virtual newline at `bytes` — but
     * only at the end of the file: in a streamed chunk the rest of that line has not arrived yet */
    if (!final_chunk) last = '\n';
    const uint64_t n_lines = (uint64_t)n_nl + (last != '\n' ? 1 : 0);
    GF_CUDA_TRY(out->nl.reserve(sizeof(unsigned long long) * (n_lines + 2)));
    unsigned long long* nl = out->nl.as<unsigned long long>();
    /* nl[0] = -1 (virtual newline before the first byte), nl[1..] = real newlines */
    const unsigned long long minus1 = ~0ull;
    GF_CUDA_TRY(cudaMemcpyAsync(nl, &minus1, sizeof(minus1), cudaMemcpyHostToDevice, st));
    k_nl_write<<<(unsigned)n_tiles, NL_THREADS, 0, st>>>(d_text, bytes, d_off, nl + 1);
    if (last != '\n') {
        const unsigned long long endpos = bytes;
        GF_CUDA_TRY(cudaMemcpyAsync(nl + 1 + n_nl, &endpos, sizeof(endpos), cudaMemcpyHostToDevice, st));
    }
    const uint64_t n_records = n_lines / 4;
    out->n_records = n_records;
    if (n_records) {
        GF_CUDA_TRY(out->s.reserve(sizeof(unsigned long long) * n_records));
        GF_CUDA_TRY(out->e.reserve(sizeof(unsigned long long) * n_records));
        GF_CUDA_TRY(out->qs.reserve(sizeof(unsigned long long) * n_records));
        unsigned int* d_flags = reinterpret_cast<unsigned int*>(d_cnt); /* reuse: [0] max_len, [1] err */
        GF_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 2 * sizeof(unsigned int), st));
        /* record i uses lines 4i..4i+3; line j spans (nl[j], nl[j+1]) in the shifted table */
        k_records<<<(unsigned)((n_records + 255) / 256), 256, 0, st>>>(nl, n_records, out->s.as<unsigned long long>(),
                                                                       out->e.as<unsigned long long>(),
                                                                       out->qs.as<unsigned long long>(), d_flags, d_flags + 1);
        unsigned int h[2] = {0, 0};
        GF_CUDA_TRY(cudaMemcpyAsync(h, d_flags, sizeof(h), cudaMemcpyDeviceToHost, st));
        GF_CUDA_TRY(cudaStreamSynchronize(st));
        out->max_len = h[0];
        if (h[1] & 1u) {
            gf_set_error("FASTQ: a quality line and its sequence line differ in length (the reference indexes qualities by "
                         "sequence position and panics when they are shorter, src/core/read.rs:349)");
            return GF_E_INVALID;
        }
    }
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}
