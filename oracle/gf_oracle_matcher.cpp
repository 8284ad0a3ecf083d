/*
 * gf_oracle_matcher.cpp — CPU restatement of the reference's `Matcher` (src/core/matcher.rs) as it is called from
 * FusionMapper::remove_alignables (src/core/fusion_mapper.rs:488-542).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see gf_oracle.h).  Parity unpinned: the reference has no test for this code and cannot
 * be built here; this file follows the cited lines literally, INCLUDING the port's bugs, because they decide the result:
 *   - every make_kmer* helper `break`s out of its loop inside the first match arm (matcher.rs:778-793, 818-834, 855-869),
 *     so it returns the 2-bit code of ONE base;
 *   - index_contig_bytes rolls the k-mer over seq[i], not seq[i + 15] (:237-252);
 *   - map_to_index adds votes with the SHADOWED enumerate index (:432-433) and its mask loop `continue`s when the key IS
 *     present (:486), then unwraps a missing key (:490-491) => panic.
 * Places where the Rust code would panic are reported as status codes instead of aborting the test process.
 */
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "gf_oracle.h"

namespace {

struct GenePos {
    int16_t contig;
    int32_t position;
};

constexpr size_t BLOOM_FILTER_LENGTH = (size_t)1 << 29; /* matcher.rs:29 */
constexpr int32_t KMER = 16;                             /* :30 */

struct RefPanic {
    int stage;
};

/* src/core/sequence.rs:52-60 */
inline uint8_t get_complement_base(uint8_t b) {
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return 'N';
    }
}
/* src/core/sequence.rs:22-50 */
std::string reverse_complement(const std::string& seq) {
    std::string out(seq.size(), 'N');
    for (size_t i = 0; i < seq.size(); i++) out[i] = (char)get_complement_base((uint8_t)seq[seq.size() - 1 - i]);
    return out;
}

/* matcher.rs:810-847 make_kmer_bytes (== make_kmer_cv :772-808 == make_kmer :849-885 on chars): the `break` inside each
 * base arm leaves the loop after the FIRST base.  `seq.get(pos..pos+16).unwrap()` panics when the slice leaves the
 * sequence: reported as RefPanic. */
uint32_t make_kmer_bytes(const uint8_t* seq, size_t len, size_t pos, bool* valid, bool slice_checked) {
    if (slice_checked && pos + (size_t)KMER > len) throw RefPanic{9};
    uint32_t kmer = 0;
    for (size_t i = 0; i < (size_t)KMER && pos + i < len; i++) {
        const uint8_t base = seq[pos + i];
        bool brk = false;
        switch (base) {
            case 'A': kmer += 0; brk = true; break;
            case 'T': kmer += 1; brk = true; break;
            case 'C': kmer += 2; brk = true; break;
            case 'G': kmer += 3; brk = true; break;
            default: *valid = false; return 0;
        }
        if (brk) break;
        if ((int32_t)i < KMER - 1) kmer <<= 2; /* never reached */
    }
    *valid = true;
    return kmer;
}
/* matcher.rs:742-770 */
inline int32_t base2num_bytes(uint8_t c) {
    switch (c) {
        case 'A': return 0;
        case 'T': return 1;
        case 'C': return 2;
        case 'G': return 3;
        default: return -1;
    }
}
/* matcher.rs:887-910 (its own gp_to_i64: a signed ADD of the position, unlike indexer.rs:697-706) */
inline GenePos shift(const GenePos& gp, int32_t i) { return GenePos{gp.contig, gp.position - i}; }
inline int64_t gp_to_i64(const GenePos& gp) { return (int64_t)((uint64_t)(int64_t)gp.contig << 32) + (int64_t)gp.position; }

}  // namespace

struct orc_matcher {
    std::map<uint32_t, std::vector<GenePos>> m_kmer_positions;
    /* vec![0; 1 << 29] (:54): calloc hands out untouched zero pages, so small tests do not pay for 512 MiB */
    uint8_t* m_bloom_filter_array = (uint8_t*)calloc(BLOOM_FILTER_LENGTH, 1);
    ~orc_matcher() { free(m_bloom_filter_array); }

    /* matcher.rs:73-88.  `0..(s.len() - 16 + 1)` is usize arithmetic: a sequence shorter than 15 bases wraps to a
     * practically endless loop (release build) or panics (debug build): reported as RefPanic stage 4. */
    void init_bloom_filter_with_seq(const std::string& s) {
        if (s.size() + 1 < (size_t)KMER) throw RefPanic{4};
        bool valid = false;
        for (size_t i = 0; i < s.size() - (size_t)KMER + 1; i++) {
            /* make_kmer on chars: skip(i).take(16) never panics */
            const uint32_t kmer = make_kmer_bytes((const uint8_t*)s.data(), s.size(), i, &valid, false);
            if (!valid) continue;
            m_bloom_filter_array[kmer >> 3] |= (uint8_t)(1 << (kmer & 0x07));
        }
    }
    /* matcher.rs:227-289.  seq is already upper-cased (:143-148). */
    void index_contig_bytes(int32_t ctg, const std::vector<uint8_t>& seq, int32_t start) {
        uint32_t kmer = 0;
        bool valid = false;
        /* seq.get(..((len as i32 - 16) as usize)).unwrap(): a negative bound wraps to a huge usize -> None -> panic */
        if ((int32_t)seq.size() - KMER < 0) throw RefPanic{1};
        const size_t n = (size_t)((int32_t)seq.size() - KMER);
        for (size_t idx = 0; idx < n; idx++) {
            const int32_t i = (int32_t)idx;
            const uint8_t base = seq[idx];
            if (valid) {
                const int32_t num = base2num_bytes(base);
                if (num < 0) {
                    valid = false;
                    continue;
                } else {
                    kmer = (kmer << 2) | (uint32_t)num;
                }
            } else {
                kmer = make_kmer_bytes(seq.data(), seq.size(), idx, &valid, true);
                if (!valid) continue;
            }
            if ((m_bloom_filter_array[kmer >> 3] & (uint8_t)(1 << (kmer & 0x07))) == 0) continue;
            m_kmer_positions[kmer].push_back(GenePos{(int16_t)ctg, i + start});
        }
    }
    /* matcher.rs:388-529.  Returns true for Some(MatchResult); throws RefPanic{2} at the `.get(&kmer).unwrap()` on a
     * missing key (:490-491). */
    bool map_to_index(const std::string& sequence) {
        std::map<int64_t, int32_t> kmer_stat;
        kmer_stat[0] = 0;
        const uint8_t* seq = (const uint8_t*)sequence.data();
        const int32_t skip_threshold = 50;
        const size_t seq_len = sequence.size();
        std::vector<uint32_t> all_kmer(seq_len, 0);
        std::vector<char> kmer_valid(seq_len, 0), skipped(seq_len, 0);
        bool valid = false;
        const int32_t n_i = (int32_t)seq_len - KMER + 1;
        if (n_i < 0) throw RefPanic{4}; /* ((seq_len as i32 - 16 + 1) as usize) wraps, the slice unwrap panics */
        for (int32_t i = 0; i < n_i; i++) {
            const uint32_t kmer = make_kmer_bytes(seq, seq_len, (size_t)i, &valid, true);
            kmer_valid[(size_t)i] = valid;
            if (!valid) continue;
            all_kmer[(size_t)i] = kmer;
            auto it = m_kmer_positions.find(kmer);
            if (it == m_kmer_positions.end()) {
                kmer_stat[0] += 1;
                continue;
            }
            if ((int32_t)it->second.size() > skip_threshold) {
                skipped[(size_t)i] = 1;
                continue;
            }
            /* `for (i, gp) in kmer_pos.iter().enumerate()` shadows the sequence offset (:432-433) */
            int32_t i2 = 0;
            for (const GenePos& gp : it->second) {
                kmer_stat[gp_to_i64(shift(gp, i2))] += 1;
                i2++;
            }
        }
        constexpr int TOP = 5;
        int64_t topgp[TOP] = {0, 0, 0, 0, 0};
        int32_t topcount[TOP] = {0, 0, 0, 0, 0};
        /* the reference iterates its hash map in unspecified order; which of several equal counts lands in the top five does
         * not change whether the function returns Some, None or panics (the mask loop below never reads topgp before it
         * panics or finishes with an all-zero mask) */
        for (const auto& kv : kmer_stat) {
            const int64_t gp = kv.first;
            const int32_t count = kv.second;
            if (gp == 0 || count <= topcount[TOP - 1]) continue;
            topgp[TOP - 1] = gp;
            topcount[TOP - 1] = count;
            for (int t = TOP - 2; t >= 0; t--) {
                if (count > topcount[t]) {
                    topcount[t + 1] = topcount[t];
                    topgp[t + 1] = topgp[t];
                    topcount[t] = count;
                    topgp[t] = gp;
                }
            }
        }
        for (int t = 0; t < TOP; t++) {
            if (topcount[t] == 0) break;
            std::vector<uint8_t> mask(seq_len, 0);
            for (int32_t i = 0; i < n_i; i++) {
                const bool v = kmer_valid[(size_t)i] != 0;
                const uint32_t kmer = all_kmer[(size_t)i];
                if (!v || m_kmer_positions.count(kmer)) continue; /* inverted test (:486) */
                /* !skipped[i] is true here (skipped is only set for present keys), so the `&&` evaluates
                 * self.m_kmer_positions.get(&kmer).unwrap() on a key that is NOT in the map: panic (:490-491) */
                throw RefPanic{2};
            }
            size_t mismatches = 0;
            for (size_t i = 0; i < seq_len; i++)
                if (mask[i] == 0) mismatches++;
            if (mismatches < 10) return true; /* Some(MatchResult) — only for seq_len < 10, which panicked above */
        }
        return false;
    }
    /* matcher.rs:662-689: Some iff either orientation returns Some */
    int do_match(const std::string& sequence) {
        const std::string rcseq = reverse_complement(sequence);
        bool a;
        try { a = map_to_index(sequence); } catch (RefPanic& p) { return p.stage == 2 ? -2 : -4; }
        bool b;
        try { b = map_to_index(rcseq); } catch (RefPanic& p) { return p.stage == 2 ? -3 : -4; }
        return (a || b) ? 1 : 0;
    }
};

extern "C" {

/* Matcher::from_ref_and_seqs (matcher.rs:44-71): init_bloom_filter over every sequence and its reverse complement
 * (:63-71), then make_index (:120-169) over the contigs in the order given (FastaReader::m_all_contigs is a BTreeMap:
 * name order; with a thread pool the push order inside a key's position list is timing dependent, this restatement uses
 * the single-thread order).  *status: 0 ok, 1 = make_index panics (a contig shorter than 16 bases, :240-243),
 * 4 = undefined (a sequence shorter than 15 bases). */
orc_matcher* orc_matcher_create(const gf_ref_contig* contigs, uint32_t n_contigs, const uint8_t* seqs, const uint64_t* seq_off,
                                uint64_t n_seqs, int* status) {
    orc_matcher* m = new orc_matcher();
    *status = 0;
    if (!m->m_bloom_filter_array) { *status = -1; return m; }
    try {
        for (uint64_t j = 0; j < n_seqs; j++) {
            const std::string s((const char*)seqs + seq_off[j], (size_t)(seq_off[j + 1] - seq_off[j]));
            m->init_bloom_filter_with_seq(s);
            m->init_bloom_filter_with_seq(reverse_complement(s));
        }
        for (uint32_t c = 0; c < n_contigs; c++) {
            std::vector<uint8_t> up((size_t)contigs[c].len);
            for (size_t k = 0; k < up.size(); k++) {
                uint8_t b = contigs[c].seq[k];
                up[k] = (b >= 'a' && b <= 'z') ? (uint8_t)(b - 32) : b; /* to_ascii_uppercase (:146) */
            }
            m->index_contig_bytes((int32_t)c, up, 0);
        }
    } catch (RefPanic& p) {
        *status = p.stage;
    }
    return m;
}
void orc_matcher_destroy(orc_matcher* m) { delete m; }
/* m_kmer_positions[k].len() for k = 0..3 (no other key can exist), bits 0..7 of m_bloom_filter_array[0], and the number of
 * OTHER non-zero bloom bytes / keys >= 4 (must be 0: checks the "degenerate" claim itself) */
void orc_matcher_counts(const orc_matcher* m, uint64_t key_positions[4], uint32_t* bloom_bits, uint64_t* other_keys) {
    for (int k = 0; k < 4; k++) key_positions[k] = 0;
    uint64_t other = 0;
    for (const auto& kv : m->m_kmer_positions) {
        if (kv.first < 4) key_positions[kv.first] = kv.second.size();
        else other++;
    }
    *bloom_bits = m->m_bloom_filter_array[0];
    *other_keys = other;
}
/* do_match: 1 = Some (the read would be removed), 0 = None, -2 / -3 = the reference panics in map_to_index of the
 * sequence / of its reverse complement, -4 = undefined (sequence shorter than 15 bases) */
int orc_matcher_do_match(orc_matcher* m, const uint8_t* seq, int32_t len) {
    return m->do_match(std::string((const char*)seq, (size_t)len));
}

/* FusionMapper::remove_alignables (fusion_mapper.rs:488-542) for one set of sequences: builds the Matcher, then calls
 * do_match per sequence in order (the `retain` closure) until one panics.  Fills the same result struct as
 * gf_alignable_filter; alignable[j] = 1 where do_match returned Some. */
int orc_remove_alignables(const gf_ref_contig* contigs, uint32_t n_contigs, const uint8_t* seqs, const uint64_t* seq_off,
                          uint64_t n_seqs, uint8_t* alignable, gf_alignable_result* res) {
    memset(res, 0, sizeof(*res));
    res->panic_seq = -1;
    for (uint64_t j = 0; j < n_seqs; j++) alignable[j] = 0;
    int status = 0;
    orc_matcher* m = orc_matcher_create(contigs, n_contigs, seqs, seq_off, n_seqs, &status);
    uint64_t other = 0;
    orc_matcher_counts(m, res->key_positions, &res->bloom_bits, &other);
    if (status != 0) {
        res->panic_stage = status;
        if (status == 4)
            for (uint64_t j = 0; j < n_seqs; j++)
                if (seq_off[j + 1] - seq_off[j] < 15) { res->panic_seq = (int64_t)j; break; }
        /* the bloom filter is complete only when the panic came from make_index */
        orc_matcher_destroy(m);
        return GF_E_REF_PANIC;
    }
    for (uint64_t j = 0; j < n_seqs; j++) {
        const int r = orc_matcher_do_match(m, seqs + seq_off[j], (int32_t)(seq_off[j + 1] - seq_off[j]));
        if (r < 0) {
            res->panic_stage = -r;
            res->panic_seq = (int64_t)j;
            orc_matcher_destroy(m);
            return GF_E_REF_PANIC;
        }
        if (r == 1) { alignable[j] = 1; res->n_removed++; }
    }
    orc_matcher_destroy(m);
    return other == 0 ? GF_OK : -100; /* -100: a key >= 4 exists, i.e. the restatement's own analysis is wrong */
}

} /* extern "C" */
