/*
 * gf_inflate.cuh — raw DEFLATE (RFC 1951) decoder for one BGZF member, written to run as ONE GPU thread per member
 * (gf_fastq.cu: k_bgzf_inflate), and compiled for the host as well so that the CPU tests can check it against zlib
 * (tests/c_driver/inflate_host.cpp, tests/test_inflate_cpu.py).
 *
 * Replaces, for blocked gzip input, the flate2 MultiGzDecoder behind FastqReader (src/core/fastq_reader.rs:39-69,149-179): a
 * .fq.gz written by bgzip / bcl2fastq is a chain of independent members of <= 64 KB text each (the size of a member stands in
 * its header), so the members of a file chunk can be inflated side by side — by thousands of GPU threads, after the COMPRESSED
 * bytes have crossed PCIe, instead of by the host cores before the text crosses it.
 *
 * A member is decoded sequentially: bit buffer, a 9-bit first-level table for the literal/length code and an 8-bit one for the
 * distance code (entry = symbol << 4 | code length; longer codes fall back to the canonical count/first-code walk), tables
 * rebuilt per dynamic block.  The tables live in caller-provided memory (shared memory on the device).  Errors are reported,
 * never ignored: bad block type, bad code lengths, a distance before the start of the member, output that does not end exactly
 * at isize, input that is not used up exactly.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GF_INF_HD __host__ __device__ __forceinline__
#else
#define GF_INF_HD inline
#endif

namespace gfinf {

constexpr int LBITS = 9, DBITS = 8; /* 2.5 KB of tables per thread: 32 threads of a block keep theirs in 80 KB of shared memory */
constexpr int MAXL = 288, MAXD = 30;

struct Tables {                     /* per decoding thread */
    uint16_t lit[1 << LBITS];       /* first-level tables: symbol << 4 | length, 0 = longer than the table's bits (or unused) */
    uint16_t dist[1 << DBITS];
    uint16_t lcount[16], dcount[16]; /* canonical code: codes per length ... */
    uint16_t lsym[MAXL], dsym[MAXD]; /* ... and the symbols in code order */
    uint8_t lens[MAXL + MAXD + 2];   /* code lengths while a block header is read */
};

enum Err { OK = 0, E_INPUT = 1, E_BTYPE = 2, E_STORED = 3, E_CODELEN = 4, E_SYMBOL = 5, E_DIST = 6, E_OUTPUT = 7, E_TRAILING = 8 };

struct Bits {
    const uint8_t* in;
    uint32_t n, pos;
    uint64_t buf;
    uint32_t cnt;
    GF_INF_HD void refill() { /* at least 32 valid bits unless the input ends (zeros are shifted in then; `over` tells) */
        while (cnt <= 56) {
            const uint64_t b = pos < n ? in[pos] : 0u;
            pos++;
            buf |= b << cnt;
            cnt += 8;
        }
    }
    GF_INF_HD uint32_t peek(uint32_t k) const { return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    GF_INF_HD void drop(uint32_t k) { buf >>= k; cnt -= k; }
    GF_INF_HD uint32_t take(uint32_t k) { const uint32_t v = peek(k); drop(k); return v; }
    /* bytes consumed so far, counting only whole bytes the decoder has really used */
    GF_INF_HD uint32_t used_bytes() const { return pos - (cnt >> 3); }
    GF_INF_HD bool over() const { return used_bytes() > n; }
};

GF_INF_HD uint32_t rev_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

/* canonical Huffman code from n code lengths: count[], sym[] and the first-level table of `bits` bits.
 * Returns 0 for a complete code, 1 for an incomplete one (allowed only for a single-code distance alphabet), -1 over-subscribed. */
GF_INF_HD int build(const uint8_t* lens, int n, uint16_t* count, uint16_t* sym, uint16_t* table, int bits) {
    uint16_t offs[16];
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[lens[s]]++;
    for (int i = 0; i < (1 << bits); i++) table[i] = 0;
    if (count[0] == n) return 0; /* no codes at all: complete in zlib's sense, any use is an error (table stays 0) */
    int left = 1;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
    }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n; s++)
        if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
    /* first-level table: walk the symbols in code order, codes are assigned in that order per length */
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= bits; l++) {
        for (int k = 0; k < count[l]; k++, idx++, code++) {
            const uint16_t entry = (uint16_t)((sym[idx] << 4) | l);
            for (uint32_t r = rev_bits(code, l); r < (1u << bits); r += 1u << l) table[r] = entry;
        }
        code <<= 1;
    }
    return left > 0 ? 1 : 0;
}

/* one symbol: first-level table, else the bit-by-bit canonical walk (puff's decode) over all lengths */
template <class BitReader>
GF_INF_HD int decode_sym(BitReader& b, const uint16_t* table, int bits, const uint16_t* count, const uint16_t* sym) {
    const uint16_t e = table[b.peek((uint32_t)bits)];
    if (e) { b.drop(e & 15u); return e >> 4; }
    int code = 0, first = 0, index = 0;
    uint64_t v = b.buf;
    for (int len = 1; len <= 15; len++) {
        code |= (int)(v & 1u);
        v >>= 1;
        const int c = count[len];
        if (code - c < first) { b.drop((uint32_t)len); return sym[index + (code - first)]; }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

GF_INF_HD void fixed_lengths(uint8_t* lens) {
    int s = 0;
    for (; s < 144; s++) lens[s] = 8;
    for (; s < 256; s++) lens[s] = 9;
    for (; s < 280; s++) lens[s] = 7;
    for (; s < 288; s++) lens[s] = 8;
    for (int d = 0; d < 30; d++) lens[288 + d] = 5;
}

/* in[0, in_len) -> out[0, out_len): both lengths are known from the member's header and trailer and must come out exactly */
GF_INF_HD int inflate_member(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& T) {
    const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    Bits b{in, in_len, 0, 0, 0};
    uint32_t op = 0;
    for (;;) {
        b.refill();
        const uint32_t last = b.take(1), type = b.take(2);
        if (type == 3) return E_BTYPE;
        if (type == 0) { /* stored: to the byte boundary, LEN, ~LEN, bytes */
            b.drop(b.cnt & 7u);
            b.refill();
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ 0xFFFFu) != nlen) return E_STORED;
            uint32_t ip = b.used_bytes(); /* byte aligned here */
            if (ip + len > in_len) return E_INPUT;
            if (op + len > out_len) return E_OUTPUT;
            for (uint32_t i = 0; i < len; i++) out[op + i] = in[ip + i];
            op += len;
            ip += len;
            b.pos = ip; b.buf = 0; b.cnt = 0;
        } else {
            if (type == 1) {
                fixed_lengths(T.lens);
                build(T.lens, 288, T.lcount, T.lsym, T.lit, LBITS);
                build(T.lens + 288, 30, T.dcount, T.dsym, T.dist, DBITS);
            } else {
                const uint32_t nlen = b.take(5) + 257, ndist = b.take(5) + 1, ncode = b.take(4) + 4;
                if (nlen > 286 || ndist > 30) return E_CODELEN;
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                for (uint32_t i = 0; i < ncode; i++) { b.refill(); cl[order[i]] = (uint8_t)b.take(3); }
                /* the code-length code is decoded through the distance tables' memory (free until the real tables are built) */
                if (build(cl, 19, T.dcount, T.dsym, T.dist, 7) != 0) return E_CODELEN;
                uint32_t i = 0;
                while (i < nlen + ndist) {
                    b.refill();
                    const int s = decode_sym(b, T.dist, 7, T.dcount, T.dsym);
                    if (s < 0) return E_CODELEN;
                    if (s < 16) { T.lens[i++] = (uint8_t)s; continue; }
                    uint32_t rep, val = 0;
                    if (s == 16) { if (i == 0) return E_CODELEN; val = T.lens[i - 1]; rep = 3 + b.take(2); }
                    else if (s == 17) rep = 3 + b.take(3);
                    else rep = 11 + b.take(7);
                    if (i + rep > nlen + ndist) return E_CODELEN;
                    while (rep--) T.lens[i++] = (uint8_t)val;
                }
                if (T.lens[256] == 0) return E_CODELEN; /* no end-of-block code */
                /* the distance lengths move out of the way before the tables that share T.lens' tail are built */
                uint8_t dl[30];
                for (uint32_t d = 0; d < 30; d++) dl[d] = d < ndist ? T.lens[nlen + d] : 0;
                for (uint32_t s = nlen; s < 288; s++) T.lens[s] = 0;
                const int rl = build(T.lens, 288, T.lcount, T.lsym, T.lit, LBITS);
                if (rl < 0 || (rl > 0 && !(T.lcount[1] == 1 && T.lcount[0] == 287))) return E_CODELEN; /* incomplete: only a single 1-bit code */
                const int rd = build(dl, 30, T.dcount, T.dsym, T.dist, DBITS);
                if (rd < 0 || (rd > 0 && !(T.dcount[1] == 1 && T.dcount[0] == 29))) return E_CODELEN;
            }
            for (;;) {
                b.refill();
                int s = decode_sym(b, T.lit, LBITS, T.lcount, T.lsym);
                if (s < 0) return E_SYMBOL;
                if (s < 256) {
                    if (op >= out_len) return E_OUTPUT;
                    out[op++] = (uint8_t)s;
                    continue;
                }
                if (s == 256) break;
                s -= 257;
                if (s >= 29) return E_SYMBOL;
                const uint32_t len = lbase[s] + b.take(lext[s]);
                b.refill();
                const int d = decode_sym(b, T.dist, DBITS, T.dcount, T.dsym);
                if (d < 0 || d >= 30) return E_SYMBOL;
                const uint32_t dist = dbase[d] + b.take(dext[d]);
                if (dist > op) return E_DIST;
                if (op + len > out_len) return E_OUTPUT;
                const uint8_t* src = out + op - dist;
                uint8_t* dst = out + op;
                for (uint32_t i = 0; i < len; i++) dst[i] = src[i]; /* overlapping on purpose when dist < len */
                op += len;
            }
        }
        if (b.over()) return E_INPUT;
        if (last) break;
    }
    if (op != out_len) return E_OUTPUT;
    if (b.used_bytes() != in_len) return E_TRAILING;
    return OK;
}

/* CRC-32 (the gzip trailer's), four bytes per step with four 256-entry tables (caller-provided: shared memory on the device) */
GF_INF_HD void crc_tables(uint32_t* t /* [4][256] */, int first, int step) {
    for (int i = first; i < 256; i += step) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        t[i] = c;
    }
}
GF_INF_HD void crc_tables_rest(uint32_t* t, int first, int step) { /* after every t[0][*] is there */
    for (int i = first; i < 256; i += step) {
        uint32_t c = t[i];
        for (int k = 1; k < 4; k++) { c = t[c & 0xFFu] ^ (c >> 8); t[k * 256 + i] = c; }
    }
}
GF_INF_HD uint32_t crc32_of(const uint8_t* p, uint32_t n, const uint32_t* t) {
    uint32_t c = 0xFFFFFFFFu, i = 0;
    for (; i < n && (((uintptr_t)(p + i)) & 3u); i++) c = t[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    for (; i + 4 <= n; i += 4) {
        c ^= *reinterpret_cast<const uint32_t*>(p + i);
        c = t[3 * 256 + (c & 0xFFu)] ^ t[2 * 256 + ((c >> 8) & 0xFFu)] ^ t[256 + ((c >> 16) & 0xFFu)] ^ t[c >> 24];
    }
    for (; i < n; i++) c = t[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return ~c;
}

/* CRC-32 of a concatenation from the CRC-32s of its parts (what zlib's crc32_combine does): crc(A B) = crc(A) * x^(8 |B|) mod P
 * ^ crc(B), on finished CRC values.  x2n[k] = x^(2^k) mod P (32 entries, made once by crc_x2n_table). */
GF_INF_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
GF_INF_HD void crc_x2n_table(uint32_t* x2n /* [32] */) {
    uint32_t p = 1u << 30; /* x^1 */
    x2n[0] = p;
    for (int n = 1; n < 32; n++) x2n[n] = p = crc_multmodp(p, p);
}
GF_INF_HD uint32_t crc_x8n(uint32_t nbytes, const uint32_t* x2n) { /* x^(8 nbytes) mod P */
    uint32_t p = 1u << 31; /* x^0 */
    uint32_t n = nbytes;
    int k = 3;
    while (n) {
        if (n & 1u) p = crc_multmodp(x2n[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

}  // namespace gfinf
