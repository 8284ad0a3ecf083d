/*
 * gf_oracle.cpp — CPU oracle: literal C++ restatement of GeneFuseRust's per-read
 * fusion-matching path.  TEST INFRASTRUCTURE, NOT PRODUCT (see gf_oracle.h).
 *
 * "Literal" = same loops, same integer widths, same iteration order as the cited
 * reference lines (paths relative to /root/reference).  Containers whose behaviour
 * is not observable (the FxHash map) are replaced by a plain open-addressed table;
 * the one container whose order IS observable (BTreeMap in Indexer::map_read,
 * src/core/indexer.rs:258,336) is std::map (ascending keys).
 *
 * parity: pinned against the reference's own known-answer vectors where they exist
 * (tests/test_oracle_kat.py); Indexer::map_read and below are "parity unpinned".
 */
#include "gf_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

/* src/core/common.rs:3-7,31-32 */
struct GenePos {
    int16_t contig;
    int32_t position;
};
constexpr int16_t DUPE_NORMAL_LEVEL = -1;
constexpr int16_t DUPE_HIGH_LEVEL = -2;

/* src/core/indexer.rs:30-38 */
constexpr uint8_t MATCH_TOP = 3, MATCH_SECOND = 2, MATCH_NONE = 1, MATCH_UNKNOWN = 0;
constexpr int32_t KMER = 16;
constexpr size_t BLOOM_FILTER_SIZE = (size_t)1 << 29;

/* open-addressed int64 -> GenePos map (stands in for HashMap<i64,GenePos,FxHasher>;
 * iteration order is not observable anywhere on the path). */
struct FlatMap {
    struct Slot {
        int64_t key;
        GenePos gp;
        uint8_t used;
    };
    std::vector<Slot> slots;
    size_t count = 0, mask = 0;
    FlatMap() { rehash(1024); }
    static inline uint64_t h(int64_t k) { return ((uint64_t)k * 0x517cc1b727220a95ULL) >> 20; }
    void rehash(size_t cap) {
        std::vector<Slot> old;
        old.swap(slots);
        slots.assign(cap, Slot{0, {0, 0}, 0});
        mask = cap - 1;
        count = 0;
        for (auto& s : old)
            if (s.used) *insert_slot(s.key) = s.gp;
    }
    GenePos* insert_slot(int64_t k) {
        size_t i = h(k) & mask;
        while (slots[i].used && slots[i].key != k) i = (i + 1) & mask;
        if (!slots[i].used) {
            slots[i].used = 1;
            slots[i].key = k;
            count++;
        }
        return &slots[i].gp;
    }
    GenePos* find(int64_t k) {
        size_t i = h(k) & mask;
        while (slots[i].used) {
            if (slots[i].key == k) return &slots[i].gp;
            i = (i + 1) & mask;
        }
        return nullptr;
    }
    const GenePos* find(int64_t k) const { return const_cast<FlatMap*>(this)->find(k); }
    void insert(int64_t k, GenePos gp) {
        if ((count + 1) * 2 > slots.size()) rehash(slots.size() * 2);
        *insert_slot(k) = gp;
    }
};

struct SeqMatch {
    int32_t seq_start, seq_end;
    GenePos start_gp;
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
    std::string s(seq.size(), 'N');
    size_t n = seq.size();
    for (size_t i = 0; i < n; i++) s[i] = (char)get_complement_base((uint8_t)seq[n - 1 - i]);
    return s;
}

/* src/core/indexer.rs:852-913 (make_kmer_bytes) == :789-850 (make_kmer_cv).  The
 * rolling update (:859-872) is written out exactly as the reference does it. */
int64_t make_kmer(const uint8_t* seq, int32_t pos, int64_t last_kmer, int32_t step) {
    int64_t kmer = 0;
    int32_t start = 0;
    if (last_kmer >= 0) {
        kmer = last_kmer;
        start = KMER - step;
        if (step == 1) kmer = (kmer & 0x3FFFFFFF) << 2;
        else if (step == 2) kmer = (kmer & 0x0FFFFFFF) << 2;
        else if (step == 3) kmer = (kmer & 0x03FFFFFF) << 2;
        else if (step == 4) kmer = (kmer & 0x00FFFFFF) << 2;
    }
    for (int32_t i = start; i < KMER; i++) {
        switch (seq[pos + i]) {
            case 'A': kmer += 0; break;
            case 'T': kmer += 1; break;
            case 'C': kmer += 2; break;
            case 'G': kmer += 3; break;
            default: return -1;
        }
        if (i < KMER - 1) kmer = kmer << 2;
    }
    return kmer;
}

/* src/core/indexer.rs:689-714 */
inline GenePos shift(const GenePos& gp, int32_t i) { return GenePos{gp.contig, gp.position - i}; }
inline int64_t gp_to_i64(const GenePos& gp) {
    int64_t ret = (int64_t)gp.contig;
    int64_t concated = (int64_t)(uint64_t)(uint32_t)gp.position; /* [position, 0] little-endian */
    return (int64_t)((uint64_t)ret << 32) | concated;
}
inline GenePos i64_to_gp(int64_t v) {
    return GenePos{(int16_t)(v >> 32), (int32_t)(v & 0x00000000FFFFFFFFLL)};
}

/* src/core/indexer.rs:716-732 */
void make_mask(uint8_t* mask, uint8_t flag, int32_t seqlen, int32_t start, int32_t kmer_size) {
    int32_t end_point = std::min(seqlen, start + kmer_size);
    for (int32_t p = start; p < end_point; p++) mask[p] = std::max(mask[p], flag);
}

/* src/core/indexer.rs:616-679 */
int segment_mask(const uint8_t* mask, int32_t seqlen, GenePos gp1, GenePos gp2, SeqMatch out[2]) {
    int n = 0;
    const int32_t ALLOWED_GAP = 10, THRESHOLD_LEN = 20;
    const int32_t targets[2] = {MATCH_TOP, MATCH_SECOND};
    const GenePos gps[2] = {gp1, gp2};
    for (int t = 0; t < 2; t++) {
        int32_t target = targets[t];
        int32_t max_start = -1, max_end = -1;
        int32_t start = 0, end = 0;
        for (;;) {
            while ((int32_t)mask[start] != target && start != seqlen - 1) start++;
            if (start >= seqlen - 1) break;
            if ((int32_t)mask[start] == target) {
                end = start + 1;
                int32_t g = 0;
                while (g < ALLOWED_GAP && (end + g) < seqlen) {
                    if ((int32_t)mask[end + g] > target) break;
                    if (end + g < seqlen && (int32_t)mask[end + g] == target) {
                        end += g + 1;
                        g = 0;
                        continue;
                    }
                    g++;
                }
                end -= 1;
                if (end - start > (max_end - max_start)) {
                    max_end = end;
                    max_start = start;
                }
                start++;
            } else {
                break;
            }
        }
        if (max_end - max_start > THRESHOLD_LEN) out[n++] = SeqMatch{max_start, max_end, gps[t]};
    }
    return n;
}

/* ---- edit distance: src/core/edit_distance.rs ---- */
/* :12-92  edit_distance_bpv (N = tmax+1 words; the char map is a 256-entry table) */
size_t edit_distance_bpv(std::vector<std::vector<uint64_t>>& cmap, const uint8_t* v, size_t vsize, size_t tmax,
                         size_t tlen, size_t N) {
    size_t d = tmax * 64 + tlen;
    uint64_t top = (uint64_t)1 << ((tlen - 1) & 63);
    uint64_t lmb = (uint64_t)1 << 63;
    std::vector<uint64_t> d0(N, 0), hp(N, 0), hn(N, 0), vp(tmax + 1, 0), vn(tmax + 1, 0);
    for (size_t i = 0; i < tmax; i++) vp[i] = ~(uint64_t)0;
    for (size_t i = 0; i < tlen; i++) vp[tmax] |= ((uint64_t)1 << (i & 63));
    for (size_t i = 0; i < vsize; i++) {
        std::vector<uint64_t>& pm = cmap[v[i]];
        if (pm.empty()) pm.assign(N, 0);
        for (size_t r = 0; r <= tmax; r++) {
            uint64_t x = pm[r];
            if (r > 0 && (hn[r - 1] & lmb) != 0) x |= 1;
            d0[r] = (((x & vp[r]) + vp[r]) ^ vp[r]) | x | vn[r];
            hp[r] = vn[r] | ~(d0[r] | vp[r]);
            hn[r] = d0[r] & vp[r];
            x = hp[r] << 1;
            if (r == 0 || (hp[r - 1] & lmb) != 0) x |= 1;
            vp[r] = (hn[r] << 1) | ~(d0[r] | x);
            if (r > 0 && (hn[r - 1] & lmb) != 0) vp[r] |= 1;
            vn[r] = d0[r] & x;
        }
        if ((hp[tmax] & top) != 0) d++;
        else if ((hn[tmax] & top) != 0) d--;
    }
    return d;
}
/* :123-162 */
size_t edit_distance_map_(const uint8_t* a, size_t asize, const uint8_t* b, size_t bsize, size_t N) {
    std::vector<std::vector<uint64_t>> cmap(256);
    size_t tmax = (asize - 1) >> 6;
    size_t tlen = asize - tmax * 64;
    for (size_t i = 0; i < tmax; i++)
        for (size_t j = 0; j < 64; j++) {
            auto& e = cmap[a[i * 64 + j]];
            if (e.empty()) e.assign(N, 0);
            e[i] |= (uint64_t)1 << j;
        }
    for (size_t i = 0; i < tlen; i++) {
        auto& e = cmap[a[tmax * 64 + i]];
        if (e.empty()) e.assign(N, 0);
        e[tmax] |= (uint64_t)1 << i;
    }
    return edit_distance_bpv(cmap, b, bsize, tmax, tlen, N);
}
/* plain DP — only used to stand in for the reference's edit_distance_dp, which
 * panics (:94-100 index into zero-length Vecs).  The panic is reported through
 * *would_panic; the value returned is the true Levenshtein distance. */
size_t levenshtein_dp(const uint8_t* a, size_t n, const uint8_t* b, size_t m) {
    std::vector<uint32_t> prev(m + 1), cur(m + 1);
    for (size_t j = 0; j <= m; j++) prev[j] = (uint32_t)j;
    for (size_t i = 1; i <= n; i++) {
        cur[0] = (uint32_t)i;
        for (size_t j = 1; j <= m; j++)
            cur[j] = std::min(std::min(prev[j], cur[j - 1]) + 1, prev[j - 1] + (a[i - 1] == b[j - 1] ? 0u : 1u));
        prev.swap(cur);
    }
    return prev[m];
}
/* :164-197 */
size_t edit_distance(const uint8_t* a, size_t asize, const uint8_t* b, size_t bsize, bool* would_panic) {
    if (asize == 0) return bsize;
    else if (bsize == 0) return asize;
    if (asize < bsize) {
        std::swap(a, b);
        std::swap(asize, bsize);
    }
    size_t vsize = ((asize - 1) >> 6) + 1;
    if (vsize > 10) {
        std::swap(a, b);
        std::swap(asize, bsize);
        vsize = ((asize - 1) >> 6) + 1;
    }
    if (vsize >= 1 && vsize <= 10) return edit_distance_map_(a, asize, b, bsize, vsize);
    if (would_panic) *would_panic = true;
    return levenshtein_dp(a, asize, b, bsize);
}

}  // namespace

struct orc_index {
    gf_params p;
    std::vector<std::string> m_fusion_seq; /* src/core/indexer.rs:77 */
    std::vector<uint8_t> m_reversed;       /* Gene::is_reversed per fusion */
    FlatMap m_kmer_pos;                    /* :74 */
    std::vector<std::vector<GenePos>> m_dupe_list; /* :76 */
    uint8_t* m_bloom_filter = nullptr;     /* :75, 512 MiB exact bitmap */
    uint64_t n_sites = 0;

    /* src/core/indexer.rs:179-241 */
    void index_contig(size_t ctg, const std::string& seq, int32_t start) {
        int64_t kmer = -1;
        const uint8_t* s = (const uint8_t*)seq.data();
        for (int32_t i = 0; i < (int32_t)seq.size() - KMER; i++) {
            kmer = make_kmer(s, i, kmer, 1);
            if (kmer < 0) continue;
            n_sites++;
            GenePos site{(int16_t)ctg, i + start};
            GenePos* found = m_kmer_pos.find(kmer);
            if (found) {
                GenePos gp = *found;
                if (gp.contig == DUPE_HIGH_LEVEL) {
                    continue;
                } else if (gp.contig == DUPE_NORMAL_LEVEL) {
                    if ((int64_t)m_dupe_list[gp.position].size() >= (int64_t)p.skip_key_dup_threshold) {
                        found->contig = DUPE_HIGH_LEVEL;
                        m_dupe_list[gp.position] = std::vector<GenePos>();
                    } else {
                        m_dupe_list[gp.position].push_back(site);
                    }
                } else {
                    std::vector<GenePos> gps;
                    gps.push_back(gp);
                    gps.push_back(site);
                    m_dupe_list.push_back(gps);
                    found->contig = DUPE_NORMAL_LEVEL;
                    found->position = (int32_t)(m_dupe_list.size() - 1);
                }
            } else {
                m_kmer_pos.insert(kmer, site);
            }
        }
    }
    /* src/core/indexer.rs:243-250 */
    void fill_bloom_filter() {
        for (auto& s : m_kmer_pos.slots)
            if (s.used) m_bloom_filter[s.key >> 3] |= (uint8_t)(1 << (s.key & 0x07));
    }

    /* src/core/indexer.rs:252-538 */
    int map_read(const uint8_t* seq, int32_t seqlen, SeqMatch out[2], uint64_t* probes1, bool* gated) const {
        std::map<int64_t, int32_t> kmer_stat;
        kmer_stat[0] = 0;
        const int32_t step = 2;
        int64_t kmer = -1;
        for (int32_t i = 0; i < seqlen - KMER + 1; i += step) {
            kmer = make_kmer(seq, i, kmer, step);
            if (probes1) (*probes1)++;
            if (kmer < 0) continue;
            int64_t pos = kmer >> 3;
            int64_t bit = kmer & 0x07;
            if ((m_bloom_filter[pos] & (uint8_t)(1 << bit)) == 0) {
                kmer_stat[0] += 1;
                continue;
            }
            const GenePos* gp = m_kmer_pos.find(kmer);
            if (gp->contig == DUPE_HIGH_LEVEL) {
                continue;
            } else if (gp->contig == DUPE_NORMAL_LEVEL) {
                const std::vector<GenePos>& dl = m_dupe_list[gp->position];
                for (size_t g = 0; g < dl.size(); g++) {
                    int64_t gplong = gp_to_i64(shift(dl[g], i));
                    kmer_stat[gplong] += 1;
                }
            } else {
                int64_t gplong = gp_to_i64(shift(*gp, i));
                kmer_stat[gplong] += 1;
            }
        }
        int64_t gp1 = 0, gp2 = 0;
        int32_t count1 = 0, count2 = 0;
        for (auto& kv : kmer_stat) {
            int64_t k = kv.first;
            int32_t v = kv.second;
            if (k != 0 && v > count1) {
                gp2 = gp1;
                count2 = count1;
                gp1 = k;
                count1 = v;
            } else if (k != 0 && v > count2) {
                gp2 = k;
                count2 = v;
            }
        }
        if (count1 * step < p.major_gene_key_requirement || count2 * step < p.minor_gene_key_requirement) return 0;
        if (gated) *gated = true;

        std::vector<uint8_t> mask((size_t)seqlen, MATCH_UNKNOWN);
        kmer = -1;
        for (int32_t i = 0; i < seqlen - KMER + 1; i++) {
            kmer = make_kmer(seq, i, kmer, 1);
            if (kmer < 0) continue;
            int64_t pos = kmer >> 3;
            int64_t bit = kmer & 0x07;
            if ((m_bloom_filter[pos] & (uint8_t)(1 << bit)) == 0) continue;
            const GenePos* gp = m_kmer_pos.find(kmer);
            if (gp->contig == DUPE_HIGH_LEVEL) {
                continue;
            } else if (gp->contig == DUPE_NORMAL_LEVEL) {
                const std::vector<GenePos>& dl = m_dupe_list[gp->position];
                for (size_t g = 0; g < dl.size(); g++) {
                    int64_t gplong = gp_to_i64(shift(dl[g], i));
                    if (std::llabs(gplong - gp1) <= 1) make_mask(mask.data(), MATCH_TOP, seqlen, i, KMER);
                    else if (std::llabs(gplong - gp2) <= 1) make_mask(mask.data(), MATCH_SECOND, seqlen, i, KMER);
                    else if (gplong == 0) make_mask(mask.data(), MATCH_NONE, seqlen, i, KMER);
                }
            } else {
                int64_t gplong = gp_to_i64(shift(*gp, i));
                if (std::llabs(gplong - gp1) <= 1) make_mask(mask.data(), MATCH_TOP, seqlen, i, KMER);
                else if (std::llabs(gplong - gp2) <= 1) make_mask(mask.data(), MATCH_SECOND, seqlen, i, KMER);
                else if (gplong == 0) make_mask(mask.data(), MATCH_NONE, seqlen, i, KMER);
            }
        }
        int32_t mismatches = 0;
        for (int32_t k = 0; k < seqlen; k++)
            if (mask[k] == MATCH_NONE || mask[k] == MATCH_UNKNOWN) mismatches++;
        if (mismatches > p.mismatch_threshold) return 0;
        return segment_mask(mask.data(), seqlen, i64_to_gp(gp1), i64_to_gp(gp2), out);
    }

    /* src/core/indexer.rs:541-608 */
    bool in_required_direction(const SeqMatch* mapping, int n) const {
        if (n < 2) return false;
        const SeqMatch* left = &mapping[0];
        const SeqMatch* right = &mapping[1];
        if (left->seq_start > right->seq_start) std::swap(left, right);
        if (left->start_gp.position > 0 && right->start_gp.position > 0) return true;
        if (left->start_gp.position < 0 && right->start_gp.position < 0) return false;
        bool lrev = m_reversed[(size_t)left->start_gp.contig] != 0;
        bool rrev = m_reversed[(size_t)right->start_gp.contig] != 0;
        if (lrev && !rrev) {
            return false;
        } else if (!lrev && rrev) {
            return true;
        } else {
            if (left->start_gp.contig < right->start_gp.contig) return true;
            /* :597-599 — the reference compares left with left (always false) */
            if (left->start_gp.contig == right->start_gp.contig &&
                std::abs(left->start_gp.position) < std::abs(left->start_gp.position))
                return true;
            else
                return false;
        }
    }

    /* src/core/fusion_mapper.rs:224-251 */
    int32_t calc_ed(const uint8_t* seq, int32_t seqlen, int32_t contig, int32_t start, int32_t end,
                    bool* would_panic) const {
        if ((start >= 0 && end <= 0) || (start <= 0 && end >= 0)) return -1;
        const std::string& fusion_seq = m_fusion_seq[(size_t)contig];
        if (std::abs(start) >= (int32_t)fusion_seq.size() || std::abs(end) >= (int32_t)fusion_seq.size()) return -2;
        std::string ss((const char*)seq, (size_t)seqlen);
        if (start < 0) {
            ss = reverse_complement(ss);
            int32_t tmp = start;
            start = -end;
            end = -tmp;
        }
        const uint8_t* ref_str = (const uint8_t*)fusion_seq.data() + start;
        size_t ref_len = (size_t)(end - start + 1);
        return (int32_t)edit_distance((const uint8_t*)ss.data(), ss.size(), ref_str, ref_len, would_panic);
    }

    /* src/core/fusion_mapper.rs:93-132 (map_read) + :154-194 (make_match) + :196-222 (calc_distance) */
    int fusion_map_read(const uint8_t* seq, int32_t seqlen, bool* mapable, gf_match* m, uint64_t* probes1,
                        bool* gated, bool* would_panic) const {
        SeqMatch mapping[2];
        int n = map_read(seq, seqlen, mapping, probes1, gated);
        if (n < 2) {
            *mapable = false;
            return 0;
        }
        *mapable = true;
        if (!in_required_direction(mapping, n)) return 0;
        SeqMatch* left = &mapping[0];
        SeqMatch* right = &mapping[1];
        if (left->seq_start > right->seq_start) std::swap(left, right);
        int32_t read_break = (left->seq_end + right->seq_start) / 2;
        left->start_gp.position += read_break;
        right->start_gp.position += read_break + 1;
        int32_t gap = right->seq_start - left->seq_end - 1;
        m->read_break = read_break;
        m->l_contig = left->start_gp.contig;
        m->l_pos = left->start_gp.position;
        m->r_contig = right->start_gp.contig;
        m->r_pos = right->start_gp.position;
        m->gap = gap;
        m->seq_len = seqlen;
        m->reversed = 0;
        int32_t left_len = read_break + 1;
        int32_t right_len = seqlen - (read_break + 1);
        m->l_dist = calc_ed(seq, left_len, m->l_contig, m->l_pos - left_len + 1, m->l_pos, would_panic);
        m->r_dist = calc_ed(seq + read_break + 1, right_len, m->r_contig, m->r_pos, m->r_pos + right_len - 1,
                            would_panic);
        /* what filter_matches would decide for this record (src/core/fusion_mapper.rs:298-377) */
        auto is_low_complexity = [](const uint8_t* s, int32_t n) { /* :559-569 + src/utils/mod.rs:48-56 */
            if (n < 20) return true;
            int32_t diff = 0;
            for (int32_t i = 0; i < n - 1; i++)
                if (s[i] != s[i + 1]) diff++;
            return diff < 7;
        };
        uint8_t ff = 0;
        if (is_low_complexity(seq, read_break + 1) || is_low_complexity(seq + read_break + 1, seqlen - (read_break + 1)))
            ff |= GF_FILTER_COMPLEXITY;                                                   /* remove_by_complexity :298-321 */
        if (m->l_dist + m->r_dist >= 5) ff |= GF_FILTER_DISTANCE;                         /* remove_by_distance :323-348 */
        if (m->l_contig == m->r_contig && std::abs(m->l_pos - m->r_pos) < p.deletion_threshold)
            ff |= GF_FILTER_INDEL;                                                        /* remove_indels :350-377 */
        m->filter_flags = ff;
        return 1;
    }
};

/* src/core/read.rs:313-440 */
static int fast_merge(const uint8_t* s1, const uint8_t* q1, int32_t len1, const uint8_t* s2raw, const uint8_t* q2raw,
                      int32_t len2, std::string* out_seq, std::string* out_qual, int32_t* out_olen,
                      int32_t* out_diff) {
    /* rc_right = m_right.reverse_complement()  (src/core/read.rs:243-261) */
    std::string str2(len2, 'N'), qual2(len2, '!');
    for (int32_t i = 0; i < len2; i++) {
        str2[i] = (char)get_complement_base(s2raw[len2 - 1 - i]);
        qual2[i] = (char)q2raw[len2 - 1 - i];
    }
    const int32_t MIN_OVERLAP = 30;
    bool overlapped = false;
    int32_t olen = MIN_OVERLAP, diff = 0, low_qual_diff = 0;
    while (olen <= std::min(len1, len2)) {
        diff = 0;
        low_qual_diff = 0;
        bool ok = true;
        int32_t offset = len1 - olen;
        for (int32_t i = 0; i < olen; i++) {
            if (s1[offset + i] != (uint8_t)str2[i]) {
                diff++;
                if ((q1[offset + i] >= '?' && (uint8_t)qual2[i] <= '0') ||
                    (q1[offset + i] <= '0' && (uint8_t)qual2[i] >= '?'))
                    low_qual_diff++;
                if (diff > low_qual_diff || low_qual_diff >= 3) {
                    ok = false;
                    break;
                }
            }
        }
        if (ok) {
            overlapped = true;
            break;
        }
        olen++;
    }
    if (!overlapped) return 0;
    int32_t offset = len1 - olen;
    std::string mseq((const char*)s1, (size_t)offset);
    mseq += str2;
    std::string mqual((const char*)q1, (size_t)offset);
    mqual += qual2;
    for (int32_t i = 0; i < olen; i++) {
        if (s1[offset + i] != (uint8_t)str2[i]) {
            if (q1[offset + i] >= '?' && (uint8_t)qual2[i] <= '0') {
                mseq[offset + i] = (char)s1[offset + i];
                mqual[offset + i] = (char)q1[offset + i];
            } else {
                mseq[offset + i] = str2[i];
                mqual[offset + i] = qual2[i];
            }
        } else {
            uint32_t q = (uint32_t)q1[offset + i] + (uint32_t)(uint8_t)qual2[i] - 33;
            if (q >= 'Z') q = 'Z';
            mqual[offset + i] = (char)q;
        }
    }
    *out_seq = mseq;
    *out_qual = mqual;
    *out_olen = olen;
    *out_diff = diff;
    return 1;
}

/* ------------------------------------------------------------------ C API */
/* ---- report stage, per clustered match: src/core/fusion_result.rs ---- */
/* :770-798 get_ref_seq */
static std::string fr_get_ref_seq(const std::string& ref_s, int32_t start, int32_t end) {
    if ((start >= 0 && end <= 0) || (start <= 0 && end >= 0)) return "";
    if (std::abs(start) >= (int32_t)ref_s.size() || std::abs(end) >= (int32_t)ref_s.size()) return "";
    size_t len = (size_t)(std::abs(end - start) + 1);
    if (start < 0) return reverse_complement(ref_s.substr((size_t)(-end), len));
    return ref_s.substr((size_t)start, len);
}
/* :324-397 FusionResult::calc_ed.  `undefined` is set where the Rust code's `as usize` casts would wrap
 * (left_len < 0 or > seq length: not reachable from make_match's read_break +-3, segments are > 20 long). */
static int32_t fr_calc_ed(const std::string& seq, int32_t m_read_break, int32_t shift, const std::string& m_left_ref,
                          const std::string& m_right_ref, int32_t* left_ed, int32_t* right_ed, bool* undefined) {
    int32_t read_break = m_read_break + shift;
    int32_t left_len = read_break + 1;
    int32_t right_len = (int32_t)seq.size() - left_len;
    if (left_len < 0 || right_len < 0) { *undefined = true; return 0; }
    std::string left_seq = seq.substr(0, (size_t)left_len);
    std::string right_seq = seq.substr((size_t)left_len, (size_t)right_len);
    /* use the sequence near the break point to adjust */
    int32_t left_comp = (int32_t)std::min(std::min(left_seq.size(), m_left_ref.size()), (size_t)20);
    int32_t right_comp = (int32_t)std::min(std::min(right_seq.size(), m_right_ref.size()), (size_t)20);
    bool wp = false;
    std::string a = left_seq.substr(left_seq.size() - (size_t)left_comp, (size_t)left_comp);
    std::string b = m_left_ref.substr(m_left_ref.size() - (size_t)left_comp, (size_t)left_comp);
    int32_t left_part_ed = (int32_t)edit_distance((const uint8_t*)a.data(), a.size(), (const uint8_t*)b.data(), b.size(), &wp);
    a = right_seq.substr(0, (size_t)right_comp);
    b = m_right_ref.substr(0, (size_t)right_comp);
    int32_t right_part_ed = (int32_t)edit_distance((const uint8_t*)a.data(), a.size(), (const uint8_t*)b.data(), b.size(), &wp);
    int32_t total_ed = left_part_ed + right_part_ed;
    /* recalculate the left and right edit distance */
    left_comp = std::min(left_len, (int32_t)m_left_ref.size());
    right_comp = std::min(right_len, (int32_t)m_right_ref.size());
    a = left_seq.substr(left_seq.size() - (size_t)left_comp, (size_t)left_comp);
    b = m_left_ref.substr(m_left_ref.size() - (size_t)left_comp, (size_t)left_comp);
    *left_ed = (int32_t)edit_distance((const uint8_t*)a.data(), a.size(), (const uint8_t*)b.data(), b.size(), &wp);
    a = right_seq.substr(0, (size_t)right_comp);
    b = m_right_ref.substr(0, (size_t)right_comp);
    *right_ed = (int32_t)edit_distance((const uint8_t*)a.data(), a.size(), (const uint8_t*)b.data(), b.size(), &wp);
    if (wp) *undefined = true;
    return total_ed;
}

extern "C" {

orc_index* orc_index_create(const gf_gene_span* genes, uint32_t n_genes, const gf_params* p) {
    orc_index* idx = new orc_index();
    idx->p = *p;
    idx->m_bloom_filter = (uint8_t*)calloc(BLOOM_FILTER_SIZE, 1);
    /* src/core/indexer.rs:122-177; chromosome lookup + slicing (:136-159) is the caller's. */
    for (uint32_t ctg = 0; ctg < n_genes; ctg++) {
        idx->m_reversed.push_back(genes[ctg].reversed);
        if (genes[ctg].len == 0) {
            idx->m_fusion_seq.push_back("");
            continue;
        }
        std::string s((const char*)genes[ctg].seq, genes[ctg].len);
        for (auto& c : s)
            if (c >= 'a' && c <= 'z') c = (char)(c - 32); /* to_uppercase (:159), ASCII */
        idx->index_contig(ctg, s, 0);
        std::string rc = reverse_complement(s);
        idx->index_contig(ctg, rc, 1 - (int32_t)s.size());
        idx->m_fusion_seq.push_back(s);
    }
    idx->fill_bloom_filter();
    return idx;
}

void orc_index_destroy(orc_index* idx) {
    if (!idx) return;
    free(idx->m_bloom_filter);
    delete idx;
}

void orc_index_counts(const orc_index* idx, uint64_t out[5]) {
    uint64_t u = 0, n = 0, h = 0;
    for (auto& s : idx->m_kmer_pos.slots)
        if (s.used) {
            if (s.gp.contig == DUPE_HIGH_LEVEL) h++;
            else if (s.gp.contig == DUPE_NORMAL_LEVEL) n++;
            else u++;
        }
    out[0] = idx->n_sites;
    out[1] = u + n + h;
    out[2] = u;
    out[3] = n;
    out[4] = h;
}

void orc_index_lookup(const orc_index* idx, const uint32_t* kmers, uint64_t n, gf_lookup* out) {
    for (uint64_t i = 0; i < n; i++) {
        gf_lookup& o = out[i];
        memset(&o, 0, sizeof(o));
        int64_t k = (int64_t)kmers[i];
        if ((idx->m_bloom_filter[k >> 3] & (1 << (k & 7))) == 0) continue;
        const GenePos* gp = idx->m_kmer_pos.find(k);
        if (gp->contig == DUPE_HIGH_LEVEL) {
            o.kind = 3;
        } else if (gp->contig == DUPE_NORMAL_LEVEL) {
            o.kind = 2;
            std::vector<GenePos> dl = idx->m_dupe_list[gp->position];
            std::sort(dl.begin(), dl.end(), [](const GenePos& a, const GenePos& b) {
                return a.contig != b.contig ? a.contig < b.contig : a.position < b.position;
            });
            o.n_sites = (int32_t)dl.size();
            for (size_t j = 0; j < dl.size() && j < 8; j++) {
                o.contig[j] = dl[j].contig;
                o.position[j] = dl[j].position;
            }
        } else {
            o.kind = 1;
            o.n_sites = 1;
            o.contig[0] = gp->contig;
            o.position[0] = gp->position;
        }
    }
}

uint64_t orc_index_keys(const orc_index* idx, uint32_t* keys, uint64_t cap) {
    uint64_t n = 0;
    for (auto& s : idx->m_kmer_pos.slots)
        if (s.used) {
            if (keys && n < cap) keys[n] = (uint32_t)s.key;
            n++;
        }
    if (keys) std::sort(keys, keys + std::min(n, cap));
    return n;
}

int orc_map_read(const orc_index* idx, const uint8_t* seq, int32_t len, orc_seqmatch out[2]) {
    SeqMatch sm[2];
    int n = idx->map_read(seq, len, sm, nullptr, nullptr);
    for (int i = 0; i < n; i++)
        out[i] = orc_seqmatch{sm[i].seq_start, sm[i].seq_end, sm[i].start_gp.contig, sm[i].start_gp.position};
    return n;
}

int orc_fusion_map_read(const orc_index* idx, const uint8_t* seq, int32_t len, int* mapable, gf_match* m) {
    bool mp = false, wp = false;
    memset(m, 0, sizeof(*m));
    m->merge_olen = -1;
    int r = idx->fusion_map_read(seq, len, &mp, m, nullptr, nullptr, &wp);
    *mapable = mp ? 1 : 0;
    return r;
}

int orc_fast_merge(const uint8_t* s1, const uint8_t* q1, int32_t len1, const uint8_t* s2, const uint8_t* q2,
                   int32_t len2, uint8_t* out_seq, uint8_t* out_qual, int32_t* out_len, int32_t* olen,
                   int32_t* diff) {
    std::string ms, mq;
    int32_t ol = 0, df = 0;
    if (!fast_merge(s1, q1, len1, s2, q2, len2, &ms, &mq, &ol, &df)) return 0;
    if (out_seq) memcpy(out_seq, ms.data(), ms.size());
    if (out_qual) memcpy(out_qual, mq.data(), mq.size());
    *out_len = (int32_t)ms.size();
    *olen = ol;
    *diff = df;
    return 1;
}

void orc_reverse_complement(const uint8_t* s, int32_t len, uint8_t* out) {
    std::string r = reverse_complement(std::string((const char*)s, (size_t)len));
    memcpy(out, r.data(), r.size());
}

int64_t orc_edit_distance(const uint8_t* a, int64_t alen, const uint8_t* b, int64_t blen) {
    bool wp = false;
    int64_t d = (int64_t)edit_distance(a, (size_t)alen, b, (size_t)blen, &wp);
    return wp ? -1000000 - d : d; /* negative = the reference would panic (DP branch); -(1e6+d) carries d */
}

int64_t orc_levenshtein_dp(const uint8_t* a, int64_t alen, const uint8_t* b, int64_t blen) {
    return (int64_t)levenshtein_dp(a, (size_t)alen, b, (size_t)blen);
}

int64_t orc_gp_to_i64(int16_t contig, int32_t position) { return gp_to_i64(GenePos{contig, position}); }
void orc_i64_to_gp(int64_t v, int16_t* contig, int32_t* position) {
    GenePos g = i64_to_gp(v);
    *contig = g.contig;
    *position = g.position;
}

int orc_segment_mask(const uint8_t* mask, int32_t seqlen, int64_t gp1, int64_t gp2, orc_seqmatch out[2]) {
    SeqMatch sm[2];
    int n = segment_mask(mask, seqlen, i64_to_gp(gp1), i64_to_gp(gp2), sm);
    for (int i = 0; i < n; i++)
        out[i] = orc_seqmatch{sm[i].seq_start, sm[i].seq_end, sm[i].start_gp.contig, sm[i].start_gp.position};
    return n;
}

int64_t orc_make_kmer(const uint8_t* seq, int32_t pos) { return make_kmer(seq, pos, -1, 1); }


/* ---- FusionMapper::add_match (src/core/fusion_mapper.rs:253-275), the per-record filters of filter_matches (:298-377, as
 * already decided in gf_match.filter_flags by orc_scan_pairs) and sort_matches (:379-385) with ReadMatch's order
 * (src/core/read_match.rs:203-229).  `names[i]` = m_read.m_name of record i (NUL-terminated).  Records are pushed in input
 * order (= the reference's push order at -t 1), flagged ones are dropped when drop_filtered != 0 (the three retain() passes),
 * every bucket is sorted with sort_by(|a, b| b.partial_cmp(a)) (stable), and the buckets are written out in index order.
 * out_index[k] = position in `in` of the k-th output record.  Returns the number written. ---- */
namespace {
struct RmLite {
    int32_t m_read_break;
    size_t seq_len;       /* m_read.m_seq.m_str.len() */
    const char* m_name;   /* m_read.m_name */
    uint64_t src;         /* index in the input */
};
/* impl PartialOrd for ReadMatch (read_match.rs:203-229): -1 Less, 0 Equal, 1 Greater */
int rm_partial_cmp(const RmLite& self, const RmLite& other) {
    if (self.m_read_break != other.m_read_break) return self.m_read_break < other.m_read_break ? -1 : 1;
    /* other.len().partial_cmp(&self.len()) */
    if (other.seq_len != self.seq_len) return other.seq_len < self.seq_len ? -1 : 1;
    const int c = strcmp(self.m_name, other.m_name); /* String's order = byte-wise, like strcmp on NUL-free names */
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
}
}  // namespace

uint64_t orc_bucket_sort(const gf_match* in, uint64_t n, uint32_t n_genes, const char* const* names, int drop_filtered,
                         uint64_t* out_index, int64_t* out_bucket) {
    std::vector<std::vector<RmLite>> fusion_matches((size_t)n_genes * n_genes); /* fusion_mapper.rs:47-49 */
    for (uint64_t i = 0; i < n; i++) {
        if (drop_filtered && in[i].filter_flags) continue;
        const int32_t index = (int32_t)n_genes * (int32_t)in[i].r_contig + (int32_t)in[i].l_contig; /* :263 */
        fusion_matches[(size_t)index].push_back(RmLite{in[i].read_break, (size_t)in[i].seq_len, names[i], i});
    }
    uint64_t w = 0;
    for (size_t b = 0; b < fusion_matches.size(); b++) {
        auto& v = fusion_matches[b];
        /* rmv.sort_by(|a, b| b.partial_cmp(a).unwrap()): a goes first iff b.partial_cmp(a) == Less; slice::sort_by is stable */
        std::stable_sort(v.begin(), v.end(), [](const RmLite& a, const RmLite& bb) { return rm_partial_cmp(bb, a) < 0; });
        for (const RmLite& r : v) {
            out_index[w] = r.src;
            if (out_bucket) out_bucket[w] = (int64_t)b;
            w++;
        }
    }
    return w;
}

static thread_local uint64_t g_counters[6];

/* src/core/pescanner.rs:427-518 (PE) / src/core/sescanner.rs:183-205 (SE) */
uint64_t orc_scan_pairs(const orc_index* idx, const gf_batch* in, gf_match* out, uint64_t cap, int threads) {
    const uint64_t PACK_SIZE = 1000; /* src/core/common.rs:23 */
    uint64_t n_packs = (in->n + PACK_SIZE - 1) / PACK_SIZE;
    if (threads < 1) threads = 1;
    std::atomic<uint64_t> next_pack(0);
    std::vector<std::vector<gf_match>> results((size_t)threads);
    std::vector<std::array<uint64_t, 6>> counters((size_t)threads, std::array<uint64_t, 6>{0, 0, 0, 0, 0, 0});
    const bool pe = in->seq2 != nullptr;

    auto worker = [&](int t) {
        std::vector<gf_match>& res = results[(size_t)t];
        std::array<uint64_t, 6>& c = counters[(size_t)t];
        auto map_one = [&](const uint8_t* seq, int32_t len, bool* mapable, gf_match* m) -> int {
            bool gated = false, wp = false;
            c[0]++;
            c[4] += (uint64_t)len;
            int r = idx->fusion_map_read(seq, len, mapable, m, &c[1], &gated, &wp);
            if (gated) c[2]++;
            if (wp) c[5]++;
            return r;
        };
        /* map a read, then its reverse complement when "mapable" (the rc-retry policy) */
        auto map_with_retry = [&](const uint8_t* seq, int32_t len, uint64_t pair_idx, uint8_t source,
                                  bool set_reversed, int16_t olen, int16_t diff) {
            gf_match m;
            memset(&m, 0, sizeof(m));
            bool mapable = false;
            int hit = map_one(seq, len, &mapable, &m);
            if (hit) {
                m.used_rc = 0;
            } else if (mapable) {
                std::string rc = reverse_complement(std::string((const char*)seq, (size_t)len));
                memset(&m, 0, sizeof(m));
                hit = map_one((const uint8_t*)rc.data(), len, &mapable, &m);
                if (hit) {
                    m.used_rc = 1;
                    m.reversed = set_reversed ? 1 : 0;
                }
            }
            if (hit) {
                m.pair_idx = pair_idx;
                m.source = source;
                m.merge_olen = olen;
                m.merge_diff = diff;
                res.push_back(m);
            }
        };
        for (;;) {
            uint64_t pk = next_pack.fetch_add(1);
            if (pk >= n_packs) break;
            uint64_t lo = pk * PACK_SIZE, hi = std::min(in->n, lo + PACK_SIZE);
            for (uint64_t p = lo; p < hi; p++) {
                const uint8_t* s1 = in->seq1 + in->off1[p];
                const uint8_t* q1 = in->qual1 + in->off1[p];
                int32_t len1 = (int32_t)(in->off1[p + 1] - in->off1[p]);
                if (!pe) {
                    map_with_retry(s1, len1, p, 1, true, -1, 0);
                    continue;
                }
                const uint8_t* s2 = in->seq2 + in->off2[p];
                const uint8_t* q2 = in->qual2 + in->off2[p];
                int32_t len2 = (int32_t)(in->off2[p + 1] - in->off2[p]);
                std::string ms, mq;
                int32_t olen = 0, diff = 0;
                if (fast_merge(s1, q1, len1, s2, q2, len2, &ms, &mq, &olen, &diff)) {
                    c[3]++;
                    map_with_retry((const uint8_t*)ms.data(), (int32_t)ms.size(), p, 0, false, (int16_t)olen,
                                   (int16_t)diff);
                    continue;
                }
                map_with_retry(s1, len1, p, 1, true, -1, 0);
                map_with_retry(s2, len2, p, 2, true, -1, 0);
            }
        }
    };
    if (threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(worker, t);
        for (auto& x : th) x.join();
    }
    std::vector<gf_match> all;
    for (auto& r : results) all.insert(all.end(), r.begin(), r.end());
    std::sort(all.begin(), all.end(), [](const gf_match& a, const gf_match& b) {
        return a.pair_idx != b.pair_idx ? a.pair_idx < b.pair_idx : a.source < b.source;
    });
    for (size_t i = 0; i < all.size() && i < cap; i++) out[i] = all[i];
    for (int k = 0; k < 6; k++) {
        g_counters[k] = 0;
        for (auto& c : counters) g_counters[k] += c[(size_t)k];
    }
    return all.size();
}

void orc_last_scan_counters(uint64_t out[6]) {
    for (int k = 0; k < 6; k++) out[k] = g_counters[k];
}


int32_t orc_get_ref_seq(const uint8_t* ref, int32_t ref_len, int32_t start, int32_t end, uint8_t* out) {
    std::string r = fr_get_ref_seq(std::string((const char*)ref, (size_t)ref_len), start, end);
    memcpy(out, r.data(), r.size());
    return (int32_t)r.size();
}

/* FusionResult::adjust_fusion_break for one match (src/core/fusion_result.rs:299-321): out = {shift, left, right} */
int orc_adjust_fusion_break(const uint8_t* seq, int32_t len, int32_t read_break, const uint8_t* left_ref, int32_t left_len,
                            const uint8_t* right_ref, int32_t right_len, int32_t out[3]) {
    std::string sq((const char*)seq, (size_t)len), lr((const char*)left_ref, (size_t)left_len),
        rr((const char*)right_ref, (size_t)right_len);
    int32_t smallest_ed = 0xFFFF, shift = 0, l = 0, r = 0;
    bool undefined = false;
    for (int32_t s = -3; s <= 3; s++) {
        int32_t left_ed = 0, right_ed = 0;
        int32_t ed = fr_calc_ed(sq, read_break, s, lr, rr, &left_ed, &right_ed, &undefined);
        if (undefined) return 1;
        if (ed < smallest_ed) {
            smallest_ed = ed;
            shift = s;
            l = left_ed;
            r = right_ed;
        }
    }
    out[0] = shift; out[1] = l; out[2] = r;
    return 0;
}

} /* extern "C" */
