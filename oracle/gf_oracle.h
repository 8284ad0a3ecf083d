/*
 * gf_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  This is a literal CPU restatement of the
 * reference's per-read fusion-matching path (GeneFuseRust, src/core/ — each
 * function in gf_oracle.cpp cites the lines it follows).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker / the CPU baseline.  The product library
 * (genefuserust_b200/libgenefuse_b200.so) never links or calls it.
 *
 * Pinning: the reference cannot be compiled here (no cargo/rustc, 11 external
 * crates).  The oracle is pinned against every known-answer vector the
 * reference's own tests hold for this path (fast_merge assert, merged fixture,
 * edit distance [0,1,90], reverse complement, gp_to_i64 round trip) — see
 * tests/test_oracle_kat.py.  Indexer::map_read, segment_mask,
 * in_required_direction, make_match and scan_pair_end are NOT pinned by any
 * reference test: for those, parity is "unpinned" and rests on this restatement.
 */
#ifndef GF_ORACLE_H
#define GF_ORACLE_H
#include <stdint.h>
#include "../include/genefuse_gpu.h" /* shares the plain record structs only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_index orc_index;

typedef struct orc_seqmatch {
    int32_t seq_start, seq_end;
    int32_t contig, position;
} orc_seqmatch;

orc_index* orc_index_create(const gf_gene_span* genes, uint32_t n_genes, const gf_params* p);
void orc_index_destroy(orc_index*);
/* counts: [0]=n_sites valid occurrences, [1]=n_keys, [2]=unique, [3]=normal, [4]=high */
void orc_index_counts(const orc_index*, uint64_t out[5]);
/* lookup like gf_index_lookup (kind 0 absent, 1 unique, 2 normal, 3 high) */
void orc_index_lookup(const orc_index*, const uint32_t* kmers, uint64_t n, gf_lookup* out);
/* dump all keys (sorted ascending) — returns count; keys may be NULL to query the size */
uint64_t orc_index_keys(const orc_index*, uint32_t* keys, uint64_t cap);

/* Indexer::map_read: returns number of SeqMatch (0..2) */
int orc_map_read(const orc_index*, const uint8_t* seq, int32_t len, orc_seqmatch out[2]);
/* FusionMapper::map_read: returns 1 when a match was made (fills m, pair_idx/source left 0) */
int orc_fusion_map_read(const orc_index*, const uint8_t* seq, int32_t len, int* mapable, gf_match* m);
/* SequenceReadPair::fast_merge: returns 1 if merged; out buffers need len1+len2 bytes */
int orc_fast_merge(const uint8_t* s1, const uint8_t* q1, int32_t len1, const uint8_t* s2, const uint8_t* q2,
                   int32_t len2, uint8_t* out_seq, uint8_t* out_qual, int32_t* out_len, int32_t* olen,
                   int32_t* diff);
void orc_reverse_complement(const uint8_t* s, int32_t len, uint8_t* out);
/* edit_distance (src/core/edit_distance.rs:164-197).  When the reference would fall into its panicking DP
 * branch (> 640 columns on both sides) the value returned is -(1000000 + d), d = true Levenshtein distance. */
int64_t orc_edit_distance(const uint8_t* a, int64_t alen, const uint8_t* b, int64_t blen);
/* textbook DP, independent of the bit-vector code above (used to cross-check it) */
int64_t orc_levenshtein_dp(const uint8_t* a, int64_t alen, const uint8_t* b, int64_t blen);
int64_t orc_gp_to_i64(int16_t contig, int32_t position);
void orc_i64_to_gp(int64_t v, int16_t* contig, int32_t* position);
int orc_segment_mask(const uint8_t* mask, int32_t seqlen, int64_t gp1, int64_t gp2, orc_seqmatch out[2]);
int64_t orc_make_kmer(const uint8_t* seq, int32_t pos);

/* scan_pair_end / scan_single_end over a whole batch, `threads` workers over
 * packs of 1000 (src/core/common.rs:23).  Output sorted by (pair_idx, source).
 * Returns number of matches (may exceed cap: then only cap are written). */
uint64_t orc_scan_pairs(const orc_index*, const gf_batch* in, gf_match* out, uint64_t cap, int threads);
/* counters of the last orc_scan_pairs on this thread: [0] sequences mapped (incl. rc retries),
 * [1] pass-1 probes, [2] sequences that passed the vote gate, [3] merged pairs, [4] seq bytes,
 * [5] edit distances that would have panicked in the reference */
void orc_last_scan_counters(uint64_t out[6]);

/* report stage: get_ref_seq (src/core/fusion_result.rs:770-798) and FusionResult::adjust_fusion_break
 * for one match (:299-397): out = {shift, m_left_distance, m_right_distance}; returns 1 where the reference's usize casts
 * would wrap (break outside the read), 0 otherwise */
int32_t orc_get_ref_seq(const uint8_t* ref, int32_t ref_len, int32_t start, int32_t end, uint8_t* out);
int orc_adjust_fusion_break(const uint8_t* seq, int32_t len, int32_t read_break, const uint8_t* left_ref, int32_t left_len,
                            const uint8_t* right_ref, int32_t right_len, int32_t out[3]);

/* add_match buckets (src/core/fusion_mapper.rs:253-275) + sort_matches (:379-385) with ReadMatch::partial_cmp
 * (src/core/read_match.rs:203-229); see gf_oracle.cpp.  out_index / out_bucket need n entries. */
uint64_t orc_bucket_sort(const gf_match* in, uint64_t n, uint32_t n_genes, const char* const* names, int drop_filtered,
                         uint64_t* out_index, int64_t* out_bucket);

/* ---- Matcher (src/core/matcher.rs) as called from FusionMapper::remove_alignables (src/core/fusion_mapper.rs:488-542);
 * gf_oracle_matcher.cpp.  Places where the Rust code panics are returned as codes. ---- */
typedef struct orc_matcher orc_matcher;
orc_matcher* orc_matcher_create(const gf_ref_contig* contigs, uint32_t n_contigs, const uint8_t* seqs, const uint64_t* seq_off,
                                uint64_t n_seqs, int* status);
void orc_matcher_destroy(orc_matcher*);
void orc_matcher_counts(const orc_matcher*, uint64_t key_positions[4], uint32_t* bloom_bits, uint64_t* other_keys);
int orc_matcher_do_match(orc_matcher*, const uint8_t* seq, int32_t len);
int orc_remove_alignables(const gf_ref_contig* contigs, uint32_t n_contigs, const uint8_t* seqs, const uint64_t* seq_off,
                          uint64_t n_seqs, uint8_t* alignable, gf_alignable_result* res);

#ifdef __cplusplus
}
#endif
#endif
