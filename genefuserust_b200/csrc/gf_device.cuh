/*
 * gf_device.cuh — device-side data layout and helpers shared by the index-build and mapping kernels.
 *
 * Index layout in HBM (replaces Indexer::{m_kmer_pos, m_dupe_list, m_bloom_filter},
 * /root/reference/src/core/indexer.rs:74-76):
 *
 *   table     open-addressed hash, buckets of 4 slots x {u32 key, u32 val} = 32 B = one HBM/L2 sector.
 *             home bucket = (key * 0x9E3779B1) >> (32 - bucket_bits); overflow goes to the next bucket.
 *             A probe that finds an EMPTY slot in a bucket stops there (no deletions ever happen).
 *   key       "plane form" of the reference's 16-mer code (A=0 T=1 C=2 G=3, indexer.rs:889-900):
 *             low 16 bits = the low code bit of base j at bit j, high 16 bits = the high code bit.  It is a
 *             bijection of the reference's 32-bit code (the hash only needs membership/values).
 *   val       [31:30] kind  0 unique  : [29] strand (1 = reverse-complement site) [28:0] goff
 *                           1 NORMAL  : [29:3] offset into `dupes`, [2:0] number of sites (2..threshold)
 *                           2 HIGH    : no payload (DUPE_HIGH_LEVEL, common.rs:32)
 *                           3 EMPTY   : 0xFFFFFFFF
 *   site      strand<<29 | goff, goff = gene_start[contig] + |position| into the padded gene arena.
 *             A reverse-strand site has position = -(forward coordinate of the k-mer's LAST base)
 *             (index_contig(rc, start = 1-len), indexer.rs:167-168,194-197).
 *   genes     upper-cased ASCII, genes laid out with >= GF_GENE_PAD bytes between them so that
 *             "goff -/+ read offset" diagonals of different genes can never collide.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GF_GENE_PAD 2304u           /* > GF_MAX_SEQ_LEN */
#define GF_EMPTY_VAL 0xFFFFFFFFu
#define GF_HASH_MULT 0x9E3779B1u
#define GF_KIND_UNIQUE 0u
#define GF_KIND_NORMAL 1u
#define GF_KIND_HIGH 2u
#define GF_SITE_STRAND (1u << 29)
#define GF_SITE_GOFF_MASK ((1u << 29) - 1u)
#define GF_MAX_GOFF (1u << 29)
#define GF_MAX_DUPES 7              /* count field is 3 bits */

struct GfDevIndex {
    const uint4* table;         /* 2 x uint4 per bucket */
    const uint32_t* dupes;      /* site codes of NORMAL keys */
    const uint8_t* gene_ascii;  /* padded arena, indexed by goff */
    const uint32_t* gene_start; /* goff of base 0 of each gene [n_genes] (ascending) */
    const uint32_t* gene_len;   /* [n_genes] */
    const uint8_t* gene_rev;    /* Gene::is_reversed() per gene */
    const uint16_t* granule_contig; /* contig owning arena granule (goff >> 11); genes are >= 2304 bytes apart, so a
                                       2048-byte granule never holds bases of two genes */
    /* L2-resident screen structures (gf_index.cu: k_gene_planes, k_window_class, filter bits in k_build_table).
     * The gene arena as interleaved bit-plane entries, 8 words (one 32-byte sector) per 32 arena positions (word w =
     * goff >> 5, bit goff & 31): {lo, hi, valid, cnt bit0, cnt bit1, cnt bit2, 0, 0}; lo / hi = code bit planes, valid =
     * upper-case ACGT, cnt = the number of sites (1..7) the forward k-mer of the window STARTING there votes for (0 when
     * that window is not an indexed site or its key is HIGH).  One 256-bit load fetches everything the diagonal
     * comparison needs about 32 gene positions. */
    const uint32_t* g_if;       /* counts of the forward k-mer */
    const uint32_t* g_ir;       /* {lo, hi, valid, ...} with the counts of the reverse-complement k-mer of that window */
    const unsigned long long* filter; /* blocked Bloom filter over all non-HIGH keys, 64-bit blocks */
    uint32_t filter_words;
    const unsigned long long* filter_multi; /* second-level filter: NORMAL (dupe) keys only */
    uint32_t filter_multi_words;
    uint32_t n_genes;
    uint32_t max_sites;         /* most sites a NORMAL key holds = max(skip_key_dup_threshold, 2) */
    uint32_t bucket_shift;      /* 32 - bucket_bits */
    uint32_t bucket_mask;
    int32_t major_req, minor_req, mismatch_thr, deletion_thr; /* gf_params */
};

/* base classification -------------------------------------------------------------------- */
/* upper-case A,C,G,T only (make_kmer_bytes, indexer.rs:888-904: anything else invalidates the k-mer) */
__device__ __forceinline__ bool gf_is_acgt_upper(uint32_t c) {
    /* bit (c-64) of {A=1, C=3, G=7, T=20} */
    return ((c & 0xE0u) == 0x40u) && ((0x0010008Au >> (c & 31u)) & 1u);
}
/* code bits for an upper-case ACGT char: A=0 T=1 C=2 G=3  ->  lo = bit 2 of c, hi = bit 1 of c */
__device__ __forceinline__ uint32_t gf_code_lo(uint32_t c) { return (c >> 2) & 1u; }
__device__ __forceinline__ uint32_t gf_code_hi(uint32_t c) { return (c >> 1) & 1u; }

/* get_complement_base, src/core/sequence.rs:52-60 (case-insensitive, everything else -> 'N') */
__device__ __forceinline__ uint8_t gf_complement_ascii(uint8_t b) {
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return 'N';
    }
}

/* 16-mer starting at s (ASCII) -> plane-form key; returns false when any base is not upper-case ACGT */
__device__ __forceinline__ bool gf_kmer_from_ascii(const uint8_t* s, uint32_t* key) {
    uint32_t lo = 0, hi = 0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t c = s[j];
        ok = ok && gf_is_acgt_upper(c);
        lo |= gf_code_lo(c) << j;
        hi |= gf_code_hi(c) << j;
    }
    *key = (hi << 16) | lo;
    return ok;
}
/* key of the reverse complement of the same window: base k of rc = complement(base 15-k); complement = code ^ 1 */
__device__ __forceinline__ uint32_t gf_key_revcomp(uint32_t key) {
    uint32_t lo = key & 0xFFFFu, hi = key >> 16;
    lo = (__brev(lo) >> 16) ^ 0xFFFFu;
    hi = __brev(hi) >> 16;
    return (hi << 16) | lo;
}
/* reference code (first base in the top bits, indexer.rs:852-913) -> plane-form key */
__device__ __forceinline__ uint32_t gf_key_from_refcode(uint32_t code) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t c = (code >> (30 - 2 * j)) & 3u;
        lo |= (c & 1u) << j;
        hi |= (c >> 1) << j;
    }
    return (hi << 16) | lo;
}

/* table probe ------------------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t gf_home_bucket(uint32_t key, uint32_t shift) { return (key * GF_HASH_MULT) >> shift; }

/* 32-byte (one sector) read-only bucket load */
__device__ __forceinline__ void gf_load_bucket(const uint4* table, uint32_t b, uint4& a, uint4& c) {
    const uint4* p = table + 2ull * b;
    /* one 256-bit load of the 32-byte bucket; L2::64B = smallest L2 prefetch size: a random probe should not drag a whole
     * 128-byte line out of HBM */
    asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(p));
}
/* match `key` inside a loaded bucket: returns val, or GF_EMPTY_VAL with *stop telling whether the probe ends */
__device__ __forceinline__ uint32_t gf_match_bucket(const uint4& a, const uint4& c, uint32_t key, bool* stop) {
    uint32_t v = GF_EMPTY_VAL;
    if (a.x == key && a.y != GF_EMPTY_VAL) v = a.y;
    if (a.z == key && a.w != GF_EMPTY_VAL) v = a.w;
    if (c.x == key && c.y != GF_EMPTY_VAL) v = c.y;
    if (c.z == key && c.w != GF_EMPTY_VAL) v = c.w;
    *stop = (v != GF_EMPTY_VAL) || a.y == GF_EMPTY_VAL || a.w == GF_EMPTY_VAL || c.y == GF_EMPTY_VAL ||
            c.w == GF_EMPTY_VAL;
    return v;
}
/* full lookup: val of `key`, GF_EMPTY_VAL when absent */
__device__ __forceinline__ uint32_t gf_table_find(const GfDevIndex& ix, uint32_t key) {
    uint32_t b = gf_home_bucket(key, ix.bucket_shift);
    for (;;) {
        uint4 a, c;
        gf_load_bucket(ix.table, b, a, c);
        bool stop;
        uint32_t v = gf_match_bucket(a, c, key, &stop);
        if (stop) return v;
        b = (b + 1) & ix.bucket_mask;
    }
}

/* membership filter (one-sided: no false negatives) -------------------------------------------------- */
/* word index + bit masks of a key: 4 "present" bits, 3 more "multi" bits set only for NORMAL (dupe) keys */
__device__ __forceinline__ uint32_t gf_filter_word(uint32_t key, uint32_t n_words) {
    return (uint32_t)(((unsigned long long)(key * GF_HASH_MULT) * n_words) >> 32);
}
/* main filter: 64-bit block = two 32-bit halves, a key sets 2 bits in each half (all 32-bit arithmetic).
 * multi filter: a second, tiny filter (16 bits per key, 4 bits per key) holding only the NORMAL (dupe) keys; it is
 * consulted only for k-mers the main filter calls present, so a false "dupe" (which would count max_sites votes)
 * needs a false positive in BOTH filters. */
__device__ __forceinline__ void gf_filter_masks(uint32_t key, uint32_t* any_lo, uint32_t* any_hi) {
    const uint32_t g1 = key * 0xC2B2AE35u, g2 = key * 0x27D4EB2Fu;
    *any_lo = (1u << (g1 >> 27)) | (1u << ((g1 >> 22) & 31u));
    *any_hi = (1u << (g2 >> 27)) | (1u << ((g2 >> 22) & 31u));
}
__device__ __forceinline__ uint32_t gf_multi_word(uint32_t key, uint32_t n_words) {
    return (uint32_t)(((unsigned long long)(key * 0x85EBCA6Bu) * n_words) >> 32);
}
__device__ __forceinline__ unsigned long long gf_multi_mask(uint32_t key) {
    const uint32_t g = key * 0x165667B1u;
    return (1ull << (g >> 26)) | (1ull << ((g >> 20) & 63u)) | (1ull << ((g >> 14) & 63u)) | (1ull << ((g >> 8) & 63u));
}
/* upper bound of the number of sites the key votes for: 0 absent/HIGH, 1 unique, max_sites dupes */
__device__ __forceinline__ uint32_t gf_filter_sites(const GfDevIndex& ix, unsigned long long w, uint32_t key, uint32_t max_sites) {
    uint32_t al, ah;
    gf_filter_masks(key, &al, &ah);
    const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
    if ((wl & al) != al || (wh & ah) != ah) return 0u;
    const unsigned long long mm = gf_multi_mask(key);
    const unsigned long long mw = __ldg(ix.filter_multi + gf_multi_word(key, ix.filter_multi_words));
    return (mw & mm) == mm ? max_sites : 1u;
}

/* level 1 only: may the key vote at all?  (A present key is then counted with the cap of the bound, min(sites, 2) = 2:
 * asking level 2 for the unique / dupe distinction would put a second, dependent L2 gather behind every hit.) */
__device__ __forceinline__ bool gf_filter_present(unsigned long long w, uint32_t key) {
    uint32_t al, ah;
    gf_filter_masks(key, &al, &ah);
    const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
    return (wl & al) == al && (wh & ah) == ah;
}

/* site decoding ---------------------------------------------------------------------------- */
/* goff -> contig: one table load */
__device__ __forceinline__ uint32_t gf_contig_of(const GfDevIndex& ix, uint32_t goff) {
    return (uint32_t)__ldg(ix.granule_contig + (goff >> 11));
}
/* site -> GenePos (contig, position) exactly as the reference stores it */
__device__ __forceinline__ void gf_site_decode(const GfDevIndex& ix, uint32_t site, int32_t* contig, int32_t* position) {
    uint32_t goff = site & GF_SITE_GOFF_MASK;
    uint32_t c = gf_contig_of(ix, goff);
    int32_t p = (int32_t)(goff - __ldg(ix.gene_start + c));
    *contig = (int32_t)c;
    *position = (site & GF_SITE_STRAND) ? -p : p;
}
/* gp_to_i64, src/core/indexer.rs:697-706 */
__device__ __forceinline__ long long gf_gp_pack(int32_t contig, int32_t position) {
    return (long long)(((unsigned long long)(long long)(int16_t)contig << 32) | (unsigned long long)(uint32_t)position);
}

__device__ __forceinline__ uint32_t gf_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t gf_lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
