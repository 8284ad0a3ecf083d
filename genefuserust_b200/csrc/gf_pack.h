/* gf_pack.h — host-side packing of ASCII reads into plane words (gf_pack.cpp); internal, not part of the ABI */
#pragma once
#include <stdint.h>

#define GF_PACK_MAX_THREADS 64

struct GfPackMate {       /* one mate of one pipeline chunk */
    const uint8_t* seq;   /* arena; read i = seq[off[i] - off_base .. off[i + 1] - off_base) */
    const uint64_t* off;  /* n + 1 offsets */
    uint64_t off_base;
    uint64_t n;
    bool mate2;           /* validity rules of R2 (case-insensitive planes + upper-case plane) instead of R1's (+ 'N' plane) */
    /* outputs, pinned host memory owned by the caller: words / xwords hold 2 * sum(ceil(len / 32)) entries, woff / xoff n */
    uint32_t* words;
    uint32_t* woff;
    uint32_t* xwords;
    uint32_t* xoff;
    uint32_t max_len;                              /* in: no read may be longer (checked by the packing threads) */
    bool compact;                                  /* in: leave out what the device can make up itself — `woff` is not written when
                                                      every read of the chunk has the same length (uniform_len), `xoff` only by
                                                      the threads that met a flagged read (xoff_written) */
    uint32_t uniform_len;                          /* out: != 0 = every read has this length (woff[i] = i * 2 * ceil(len / 32)) */
    uint8_t xoff_written[GF_PACK_MAX_THREADS];     /* out (compact): thread t wrote its part of `xoff` (reads n t / nt .. n (t + 1) / nt) */
    uint32_t bad_offsets;                          /* out: != 0 = offsets that do not ascend, or a read longer than max_len:
                                                      nothing of this chunk may be used */
    uint64_t n_words;                              /* entries of `words` in use */
    int n_threads;
    uint64_t xregion_start[GF_PACK_MAX_THREADS];   /* per packing thread: its part of `xwords` (entries) */
    uint64_t xregion_used[GF_PACK_MAX_THREADS];
};

bool gf_pack_available();                          /* AVX-512BW present and GF_HOST_PACK != 0 */
bool gf_pack_forced();                             /* GF_HOST_PACK=1: every chunk is packed (tests) */
int gf_pack_threads();                             /* GF_PACK_THREADS or the hardware threads */
void gf_pack_chunk(GfPackMate* mates, int n_mates);
/* the same in two halves: gf_pack_start returns at once (the packing threads work; `mates` must stay where it is), gf_pack_wait
 * returns when they are done, with the milliseconds they took.  One job at a time in the process: a second caller waits, or
 * (wait_if_busy = false) is told so and uploads its chunk as ASCII instead. */
bool gf_pack_start(GfPackMate* mates, int n_mates, bool check_only, bool wait_if_busy); /* false: busy and !wait_if_busy */
float gf_pack_wait();
/* only the offset check of gf_pack_chunk (sets bad_offsets), by the same threads */
void gf_pack_check_offsets(GfPackMate* mates, int n_mates);
