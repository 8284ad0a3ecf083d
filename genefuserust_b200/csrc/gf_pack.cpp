/*
 * gf_pack.cpp — host side of the packed upload of gf_map_pairs (no CUDA here).
 *
 * End to end the mapping call is bound by the host -> device copy: 316 bytes per 2x150 pair at ~50 GB/s, while the device
 * needs 0.5 ns per pair.  The kernels only ever look at bit-planes of the reads (csrc/gf_screen_tpp.cuh), so when the arenas
 * live in pinned host memory the planes are built HERE, by all host cores at memory speed (AVX-512BW: a 64-byte load and two
 * test-into-mask instructions give 64 plane bits each), and only they cross PCIe: 2 bits per base + 8 bytes per read
 * (~112 bytes per pair with the offsets).  The ASCII stays where it is; the handful of reads that survive the screen are
 * fetched from the pinned arenas by k_exact / k_verify on demand, like the quality bytes of fast_merge.
 *
 * Per read of mate m, forward orientation, bit j of word k = base 32 k + j:
 *   words[woff[i] ..]           nw = ceil(len / 32) words of `lo` (bit 2 of the ASCII byte), then nw words of `hi` (bit 1),
 *                               both cleared where the base is not valid (A0 T1 C2 G3 = hi:lo, src/core/indexer.rs:888-904)
 *   xoff[i] == 0                every base is upper-case ACGT: the validity planes are all ones, nothing else is sent
 *   xwords[xoff[i] - 1 ..]      otherwise nw words of `valid` then nw words of `aux`:
 *                                 mate 1: valid = upper-case ACGT,       aux = the byte is 'N'   (fast_merge, read.rs:339-367)
 *                                 mate 2: valid = ACGT in either case,   aux = upper-case ACGT   (reverse_complement keeps the
 *                                         case-insensitive complement, sequence.rs:52-60; map_read wants upper case)
 * which is exactly what convert_r1 / convert_r2_rc compute on the device from the ASCII (the device reverses mate 2 itself).
 * tests/test_gpu_parity.py::test_packed_upload_parity runs both paths on ragged reads with N / lower case / IUPAC bytes.
 *
 * The packer is compiled for AVX-512BW by a function attribute and only used when the CPU has it (gf_pack_available); without
 * it gf_map_pairs keeps uploading the ASCII arenas.
 */
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "gf_pack.h"

namespace {

int g_prefetch = 8192; /* bytes ahead of the read being packed (GF_PACK_PREFETCH; measured 0 .. 16384 on a B200 host: 12 threads
                          pack 10 M pairs in 52.7 ms without, 46.3 ms at 512, 32.4 ms at 8192: the arenas are 4 KB pages in a guest,
                          the hardware prefetchers stop at every page end) */

/* persistent workers: a pipeline chunk is packed in a millisecond or two, thread start-up would cost as much.  The caller
 * starts a job and does its own work (issuing the copies and launches of the chunk before) until it needs the result. */
class Pool {
public:
    explicit Pool(int n) : n_(n) {
        for (int t = 0; t < n_; t++) th_.emplace_back([this, t] { worker(t); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_++;
            agen_.store(gen_, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return n_; }
    /* fn(tid) on every worker; returns at once */
    void start(const std::function<void(int)>* fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = fn;
            left_ = n_;
            aleft_.store(left_, std::memory_order_release);
            gen_++;
            agen_.store(gen_, std::memory_order_release);
        }
        cv_.notify_all();
    }
    /* until the job is done; returns the steady-clock time at which the last worker finished */
    std::chrono::steady_clock::time_point wait() {
        for (int spin = 0; spin < 20000 && aleft_.load(std::memory_order_acquire) != 0; spin++) _mm_pause();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return left_ == 0; });
        fn_ = nullptr;
        return t_done_;
    }
    /* all workers of the running job meet here */
    void barrier() {
        const int g = bar_gen_.load(std::memory_order_acquire);
        if (bar_cnt_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            bar_cnt_.store(0, std::memory_order_relaxed);
            bar_gen_.store(g + 1, std::memory_order_release);
        } else {
            while (bar_gen_.load(std::memory_order_acquire) == g) _mm_pause();
        }
    }

private:
    void worker(int tid) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            /* a pipeline hands out a chunk every millisecond or so: spin for a moment before going to sleep */
            for (int spin = 0; spin < 20000 && agen_.load(std::memory_order_acquire) == seen; spin++) _mm_pause();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(tid);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --left_;
                if (left_ == 0) t_done_ = std::chrono::steady_clock::now();
                aleft_.store(left_, std::memory_order_release);
                if (left_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    uint64_t gen_ = 0;
    int left_ = 0;
    bool stop_ = false;
    std::chrono::steady_clock::time_point t_done_{};
    std::atomic<int> bar_cnt_{0}, bar_gen_{0};
    std::atomic<int> aleft_{0};
    std::atomic<uint64_t> agen_{0}; /* copy of gen_ the workers can poll without the lock */
};

std::mutex g_pool_mu; /* one job at a time (handles on several devices share the host cores) */
Pool* g_pool = nullptr;

int want_threads() {
    if (const char* e = getenv("GF_PACK_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= GF_PACK_MAX_THREADS) return v;
    }
    /* three quarters of the hardware threads: with the prefetches the packers reach the memory bandwidth of the box before
     * they run out of cores (measured on a 16-core B200 host: 12 threads 38.8 ms per 10 M pairs, 16 threads 37 - 40 ms), and
     * the thread that issues the copies and launches, and the CUDA driver's own, need somewhere to run */
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return (int)std::min<unsigned>(std::max(1u, hc - hc / 4), GF_PACK_MAX_THREADS);
}

/* one read (<= 1024 bases) -> plane words.  Returns true when some base is not upper-case ACGT (then xv / xa were written). */
__attribute__((target("avx512f,avx512bw"))) inline bool pack_read(const uint8_t* s, uint32_t len, bool mate2, uint32_t* lo,
                                                                  uint32_t* hi, uint32_t* xv, uint32_t* xa) {
    const __m512i b4 = _mm512_set1_epi8(4), b2 = _mm512_set1_epi8(2), m7 = _mm512_set1_epi8(7), up = _mm512_set1_epi8((char)0xDF);
    /* the letter a valid base must be, by the low 3 bits of the byte: 1 A, 3 C, 4 T, 7 G (either case); 0xFF elsewhere: a byte
     * with those low bits never equals it */
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i cn = _mm512_set1_epi8('N');
    uint64_t L[16], H[16], V[16], A[16]; /* 64 bases per entry; the words are copied out at the end */
    bool flagged = false;
    uint32_t nb = 0;
    for (uint32_t p = 0; p < len; p += 64, nb++) {
        const uint32_t rem = len - p;
        const __mmask64 km = rem >= 64 ? ~0ull : ((1ull << rem) - 1ull);
        const __m512i x = _mm512_maskz_loadu_epi8(km, s + p);
        const __m512i e = _mm512_shuffle_epi8(lut, _mm512_and_si512(x, m7));
        const uint64_t vcs = _mm512_mask_cmpeq_epi8_mask(km, x, e);
        const uint64_t l = _mm512_test_epi8_mask(x, b4), h = _mm512_test_epi8_mask(x, b2); /* bytes beyond the read were loaded as 0 */
        if (__builtin_expect(vcs == km, 1)) { /* every byte of the block is upper-case ACGT */
            L[nb] = l; H[nb] = h; V[nb] = km; A[nb] = mate2 ? km : 0ull;
        } else {
            flagged = true;
            const uint64_t v = mate2 ? (uint64_t)_mm512_mask_cmpeq_epi8_mask(km, _mm512_and_si512(x, up), e) : vcs;
            L[nb] = l & v; H[nb] = h & v; V[nb] = v;
            A[nb] = mate2 ? vcs : (uint64_t)_mm512_mask_cmpeq_epi8_mask(km, x, cn);
        }
    }
    const uint32_t bytes = 4 * ((len + 31) >> 5);
    memcpy(lo, L, bytes);
    memcpy(hi, H, bytes);
    if (flagged) { memcpy(xv, V, bytes); memcpy(xa, A, bytes); }
    return flagged;
}

/* The same for the read that has nothing to flag — every base upper-case ACGT, at most 256 bases — which is nearly every read:
 * the mask words go from the registers straight to their place (8 bytes per 64 bases, 4 for an odd last word; no staging
 * arrays, no variable-length copies), nothing is masked but the last block, and the first byte that is not ACGT sends the whole
 * read through pack_read instead (false).  `overread`: 64 bytes may be loaded from any block start (the arena goes on behind
 * the read); otherwise the last block is a masked load. */
__attribute__((target("avx512f,avx512bw"))) inline bool pack_read_clean(const uint8_t* s, uint32_t len, uint32_t nw, bool overread,
                                                                        uint32_t* lo, uint32_t* hi) {
    const __m512i b4 = _mm512_set1_epi8(4), b2 = _mm512_set1_epi8(2), m7 = _mm512_set1_epi8(7);
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1));
    uint32_t p = 0, w = 0;
    for (; p + 64 <= len; p += 64, w += 2) { /* whole blocks */
        const __m512i x = _mm512_loadu_si512((const void*)(s + p));
        const __m512i e = _mm512_shuffle_epi8(lut, _mm512_and_si512(x, m7));
        if (_mm512_cmpneq_epi8_mask(x, e)) return false;
        const uint64_t l = _mm512_test_epi8_mask(x, b4), h = _mm512_test_epi8_mask(x, b2);
        memcpy(lo + w, &l, 8);
        memcpy(hi + w, &h, 8);
    }
    if (p < len) {
        const uint32_t rem = len - p;
        const __mmask64 km = (1ull << rem) - 1ull;
        const __m512i x = overread ? _mm512_loadu_si512((const void*)(s + p)) : _mm512_maskz_loadu_epi8(km, s + p);
        const __m512i e = _mm512_shuffle_epi8(lut, _mm512_and_si512(x, m7));
        if (_mm512_mask_cmpneq_epi8_mask(km, x, e)) return false;
        const uint64_t l = _mm512_test_epi8_mask(x, b4) & km, h = _mm512_test_epi8_mask(x, b2) & km;
        if (nw - w >= 2) {
            memcpy(lo + w, &l, 8);
            memcpy(hi + w, &h, 8);
        } else {
            lo[w] = (uint32_t)l;
            hi[w] = (uint32_t)h;
        }
    }
    return true;
}

/* The plane words of one packing thread are one sequential stream.  They are collected in a small buffer (L1) and leave it as
 * whole 64-byte lines with non-temporal stores: the packers are bound by the host's memory system, and an ordinary store costs
 * a read of the line (for ownership) on top of the write; the copy engine reads the words from memory anyway. */
struct WordStream {
    static constexpr uint32_t CAP = 1024; /* words collected before they are flushed (a read adds <= 64) */
    uint32_t* dst;   /* 64-byte aligned: where stage[0] belongs */
    uint32_t fill;   /* words in stage, counting the `head` words in front */
    uint32_t head;   /* words of the first line that belong to the thread before this one: never written */
    bool nt;
    alignas(64) uint32_t stage[CAP + 64 + 16];
    __attribute__((target("avx512f,avx512bw"))) void open(uint32_t* first, bool use_nt) {
        head = (uint32_t)(((uintptr_t)first & 63u) >> 2);
        dst = first - head;
        fill = head;
        nt = use_nt;
    }
    uint32_t* cursor() { return stage + fill; }
    __attribute__((target("avx512f,avx512bw"))) void flush(bool final) {
        uint32_t done = 0;
        if (head) {
            if (fill < 16 && !final) return;
            const uint32_t n0 = fill < 16 ? fill : 16;
            memcpy(dst + head, stage + head, 4 * (size_t)(n0 - head));
            done = n0;
            head = 0;
        }
        if (nt) for (; done + 16 <= fill; done += 16) _mm512_stream_si512((__m512i*)(dst + done), _mm512_load_si512((const void*)(stage + done)));
        else for (; done + 16 <= fill; done += 16) _mm512_store_si512((void*)(dst + done), _mm512_load_si512((const void*)(stage + done)));
        if (final && fill > done) {
            memcpy(dst + done, stage + done, 4 * (size_t)(fill - done));
            done = fill;
        }
        const uint32_t rem = fill - done; /* < 16 */
        if (rem) memcpy(stage, stage + done, 4 * (size_t)rem);
        dst += done;
        fill = rem;
    }
};
bool g_nt_stores = true; /* GF_PACK_NT=0: ordinary stores (experiments) */

__attribute__((target("avx512f,avx512bw"))) void pack_range(GfPackMate* m, int t, uint64_t a, uint64_t b, uint64_t word_base, uint64_t x_base,
                                                            uint64_t* x_used, uint32_t uniform_len) {
    uint64_t P = word_base, X = x_base;
    const uint64_t end_all = m->off[m->n] - m->off_base; /* the chunk's bytes: nothing is loaded from beyond them */
    const uint8_t* const seq = m->seq - m->off_base;
    const uint32_t pf = (uint32_t)g_prefetch;
    const bool write_woff = !(m->compact && uniform_len);
    bool write_xoff = !m->compact;
    WordStream ws;
    ws.open(m->words + 2 * word_base, g_nt_stores);
    uint64_t o = a < b ? m->off[a] : 0;
    for (uint64_t i = a; i < b; i++) {
        const uint64_t o_next = m->off[i + 1];
        const uint32_t len = (uint32_t)(o_next - o), nw = (len + 31) >> 5;
        const uint8_t* s = seq + o;
        /* the arenas stream through once: ask for the lines a few reads ahead (one core alone does not keep enough misses
         * in flight to reach its share of the memory bandwidth) */
        if (pf) {
            _mm_prefetch((const char*)(s + pf), _MM_HINT_T0);
            _mm_prefetch((const char*)(s + pf + 64), _MM_HINT_T0);
            _mm_prefetch((const char*)(s + pf + 128), _MM_HINT_T0);
        }
        uint32_t* lo = ws.cursor();
        if (write_woff) m->woff[i] = (uint32_t)(2 * P);
        uint32_t xo = 0;
        const bool overread = (o - m->off_base) + (uint64_t)(len & ~63u) + 64 <= end_all;
        if (!pack_read_clean(s, len, nw, overread, lo, lo + nw)) {
            uint32_t* xv = m->xwords + X;
            if (pack_read(s, len, m->mate2, lo, lo + nw, xv, xv + nw)) {
                xo = (uint32_t)(X + 1);
                X += 2 * nw;
                if (!write_xoff) { /* the first flagged read of this thread: from here on its part of xoff is written */
                    memset(m->xoff + a, 0, sizeof(uint32_t) * (size_t)(i - a));
                    write_xoff = true;
                }
            }
        }
        if (write_xoff) m->xoff[i] = xo;
        ws.fill += 2 * nw;
        if (ws.fill >= WordStream::CAP) ws.flush(false);
        P += nw;
        o = o_next;
    }
    ws.flush(true);
    _mm_sfence(); /* the non-temporal stores are visible before the job counts as done */
    m->xoff_written[t] = write_xoff ? 1 : 0;
    *x_used = X - x_base;
}

}  // namespace

bool gf_pack_available() {
    static const bool cpu_ok = [] {
        __builtin_cpu_init();
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    }();
    const char* e = getenv("GF_HOST_PACK"); /* read per call: tests switch it */
    if (!cpu_ok || (e && atoi(e) == 0)) return false;
    /* Packing trades host memory bandwidth for PCIe bandwidth: it reads every byte once to save two thirds of the copy.  That
     * pays when this process has cores (and their share of the memory system) to spare — 12 threads on a 16-core single-GPU
     * host: 40 ms instead of 64 per 10 M pairs; 9 threads per rank on a 24-core 2-GPU host: 56 instead of 64 — and it does
     * NOT when many ranks share few cores and one memory system: 8 ranks with 3 threads each on a 32-core 8-GPU host, where
     * the eight copy engines already saturate the host (237 GB/s aggregate): 178 ms instead of 146; 4 ranks with 6 threads each
     * (final packer, two-ended upload): 83 ms instead of 67.  Fewer than 8 packing
     * threads (GF_PACK_THREADS, or 3/4 of the hardware threads): the ASCII is copied as it is.  GF_HOST_PACK=1 forces packing. */
    int min_threads = 8;
    if (const char* m = getenv("GF_PACK_MIN_THREADS")) { const int v = atoi(m); if (v >= 1) min_threads = v; } /* experiments */
    return (e && atoi(e) == 1) || want_threads() >= min_threads;
}

bool gf_pack_forced() {
    const char* e = getenv("GF_HOST_PACK");
    return e && atoi(e) == 1;
}

int gf_pack_threads() { return want_threads(); }

/* the job in flight (one at a time: g_pool_mu is held from gf_pack_start to gf_pack_wait) */
struct PackJob {
    GfPackMate* mates = nullptr;
    int n_mates = 0, nt = 1, nv = 1;
    std::atomic<int> next_a{0}, next_b{0};
    bool check_only = false;
    std::vector<uint64_t> sums;
    std::atomic<uint32_t> bad[2], ragged[2];
    std::function<void(int)> fn;
    std::chrono::steady_clock::time_point t_start{};
};
static PackJob g_job;

/* A chunk is cut into nv parts per mate (the output layout — where a part's plane words and exception words start — depends
 * on the parts, not on who packs them) and the workers take part after part from a counter.  nv = the number of threads: more,
 * smaller parts were measured (GF_PACK_PARTS) and are slower, 12 / 24 / 48 / 64 parts with 12 threads: 35.6 / 37.9 / 41.2 /
 * 43.2 ms per 10 M pairs — the memory system likes few long streams. */
static void pack_worker(int /*t*/) {
    PackJob& J = g_job;
    const int nv = J.nv;
    /* pass A: plane words of every part (from the offsets alone) + the offset check */
    for (int v = J.next_a.fetch_add(1, std::memory_order_relaxed); v < nv * J.n_mates; v = J.next_a.fetch_add(1, std::memory_order_relaxed)) {
        const int k = v / nv, vi = v % nv;
        const GfPackMate& m = J.mates[k];
        const uint64_t a = m.n * (uint64_t)vi / nv, b = m.n * (uint64_t)(vi + 1) / nv;
        uint64_t s = 0, bad = 0, differ = 0;
        const uint64_t len0 = m.n ? m.off[1] - m.off[0] : 0; /* the chunk's first read: is every read as long? */
        for (uint64_t i = a; i < b; i++) {
            const uint64_t len = m.off[i + 1] - m.off[i]; /* wraps to a huge value when the offsets descend */
            bad |= len > m.max_len;
            differ |= len ^ len0;
            s += (len + 31) >> 5;
        }
        J.sums[(size_t)k * nv + vi] = s;
        if (bad) J.bad[k].store(1, std::memory_order_relaxed);
        if (differ || len0 == 0) J.ragged[k].store(1, std::memory_order_relaxed);
    }
    if (J.check_only) return;
    g_pool->barrier();
    for (int k = 0; k < J.n_mates; k++)
        if (J.bad[k].load(std::memory_order_relaxed)) return; /* every thread sees the same flags after the barrier */
    /* pass B: pack.  A part's exception words start where its plane words would if every read had them: the regions
     * never overlap, and only their used parts are copied to the device */
    for (int v = J.next_b.fetch_add(1, std::memory_order_relaxed); v < nv * J.n_mates; v = J.next_b.fetch_add(1, std::memory_order_relaxed)) {
        const int k = v / nv, vi = v % nv;
        GfPackMate& m = J.mates[k];
        const uint64_t a = m.n * (uint64_t)vi / nv, b = m.n * (uint64_t)(vi + 1) / nv;
        uint64_t base = 0;
        for (int u = 0; u < vi; u++) base += J.sums[(size_t)k * nv + u];
        m.xregion_start[vi] = 2 * base;
        const uint32_t uniform_len = J.ragged[k].load(std::memory_order_relaxed) ? 0u : (uint32_t)(m.off[1] - m.off[0]);
        if (vi == 0) m.uniform_len = uniform_len;
        pack_range(&m, vi, a, b, base, 2 * base, &m.xregion_used[vi], uniform_len);
        if (vi == nv - 1) m.n_words = 2 * (base + J.sums[(size_t)k * nv + vi]);
    }
}

bool gf_pack_start(GfPackMate* mates, int n_mates, bool check_only, bool wait_if_busy) { /* n_mates <= 2 */
    if (wait_if_busy) g_pool_mu.lock();
    else if (!g_pool_mu.try_lock()) return false; /* another handle's chunk is being packed */
    g_prefetch = 8192;
    if (const char* e = getenv("GF_PACK_PREFETCH")) { const int v = atoi(e); if (v >= 0 && v <= (1 << 20)) g_prefetch = v; }
    const int nt = want_threads();
    if (!g_pool || g_pool->size() != nt) {
        delete g_pool;
        g_pool = new Pool(nt);
    }
    PackJob& J = g_job;
    J.mates = mates;
    J.n_mates = n_mates;
    J.nt = nt;
    J.check_only = check_only;
    J.nv = nt;
    if (const char* e = getenv("GF_PACK_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= GF_PACK_MAX_THREADS) J.nv = v; } /* experiments */
    J.next_a.store(0);
    J.next_b.store(0);
    J.sums.assign((size_t)n_mates * J.nv, 0);
    J.bad[0].store(0);
    J.bad[1].store(0);
    J.ragged[0].store(0);
    J.ragged[1].store(0);
    { const char* e = getenv("GF_PACK_NT"); g_nt_stores = !(e && atoi(e) == 0); }
    for (int k = 0; k < n_mates; k++) mates[k].n_threads = J.nv; /* (parts: what the output layout depends on) */
    if (!J.fn) J.fn = pack_worker;
    J.t_start = std::chrono::steady_clock::now();
    g_pool->start(&J.fn);
    return true;
}

float gf_pack_wait() {
    const auto t_done = g_pool->wait();
    PackJob& J = g_job;
    for (int k = 0; k < J.n_mates; k++) J.mates[k].bad_offsets = J.bad[k].load();
    const float ms = std::chrono::duration<float, std::milli>(t_done - J.t_start).count();
    g_pool_mu.unlock();
    return ms;
}

void gf_pack_chunk(GfPackMate* mates, int n_mates) {
    gf_pack_start(mates, n_mates, false, true);
    gf_pack_wait();
}
void gf_pack_check_offsets(GfPackMate* mates, int n_mates) {
    gf_pack_start(mates, n_mates, true, true);
    gf_pack_wait();
}
