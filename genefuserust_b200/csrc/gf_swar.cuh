/*
 * gf_swar.cuh — 16 ASCII bases -> 16 plane bits with a handful of SWAR operations (no per-base work).  Shared by the read
 * converters of the screen (gf_screen_tpp.cuh) and the reference scan of the Matcher pass (gf_matcher.cu).
 * tests/test_swar_bits.py models expect4 / block16 on the CPU with the constants read from THIS file.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swar {

__device__ __forceinline__ uint32_t zero_bytes(uint32_t y) { /* bit 7 of every byte that is 0 */
    return ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;
}

/* The converters read one aligned 32-byte sector (two 16-byte blocks) per step and turn each block into 16 plane bits.
 * Everything happens in the NIBBLE domain: z = the low nibbles of 8 bytes (nibble 2 i = byte i of the first word, nibble
 * 2 i + 1 = byte i of the second), y = their high nibbles in the same order.
 *   - the code bits (bit 2 / bit 1 of the ASCII byte = A0 T1 C2 G3) of 8 bytes are gathered by ONE multiply:
 *     (z & 0x44444444) * 0x00408102 has the eight bit-2 values in base order in its top byte (0x00810204 for bit 1); the
 *     partial products never collide, so there are no carries;
 *   - validity, with L / H the nibbles of a byte: upper-case ACGT <=> H = 010?, L3 = 0, H0 = L2 & ~L1 (only 'T' = 0x54 has the
 *     0x10 bit) and L0 = ~H0 (A 0x41, C 0x43, G 0x47: L = 0001, 0011, 0111, H0 = 0; T: L = 0100, H0 = 1).  The bits are lined up
 *     at the TOP of the nibble, which takes left shifts only — multiplies on the FMA pipe; the converters are bound by the ALU
 *     pipe (logic / shift / permute: one warp instruction per 2 cycles and scheduler), not by issue.  A block whose 16 bytes
 *     are all upper-case ACGT takes the fast path (valid = 0xFFFF);
 *   - misalignment a (0..31) of the read w.r.t. the sectors is removed in the bit domain: planes are built at bit position
 *     u = p + a and aligned word w = funnelshift(U[w], U[w+1], a). */
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) { /* (a & m) | (b & ~m) as ONE logic operation */
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}
constexpr uint32_t NIB3 = 0x88888888u;     /* bit 3 of every nibble */
constexpr uint32_t GATHER3 = 0x00204081u;  /* (x & NIB3) * GATHER3: the eight nibble-bit-3 values in base order in the top byte */
/* bit 3 of every nibble: the byte is NOT consistent with an ACGT letter in its low five bits (H0, L) */
__device__ __forceinline__ uint32_t nib_code_bad(uint32_t z, uint32_t y) {
    const uint32_t l2 = z << 1, l1 = z << 2, l0 = z << 3, h0 = y << 3;
    return (h0 ^ (l2 & ~l1)) | ~(l0 ^ h0);
}
/* One 16-byte block -> 16 plane bits in the LOW half of each result (the upper halves are garbage):
 *   lo / hi = code bits, v = valid (ACGT; either case when CI), ex = !CI: the byte is 'N';  CI: valid AND upper case */
template <bool CI>
__device__ __forceinline__ void block16(const uint4& x, uint32_t* lo, uint32_t* hi, uint32_t* v, uint32_t* ex) {
    const uint32_t z0 = bitsel(x.x, x.y << 4, 0x0F0F0F0Fu), z1 = bitsel(x.z, x.w << 4, 0x0F0F0F0Fu);
    const uint32_t y0 = bitsel(x.x >> 4, x.y, 0x0F0F0F0Fu), y1 = bitsel(x.z >> 4, x.w, 0x0F0F0F0Fu);
    uint32_t l = __byte_perm((z0 & 0x44444444u) * 0x00408102u, (z1 & 0x44444444u) * 0x00408102u, 0x7373u);
    uint32_t h = __byte_perm((z0 & 0x22222222u) * 0x00810204u, (z1 & 0x22222222u) * 0x00810204u, 0x7373u);
    /* all 16 bytes upper-case ACGT?  (code consistency or L3 at bit 3; the high nibble must be 0100 or 0101) */
    const uint32_t c0 = nib_code_bad(z0, y0) | z0, c1 = nib_code_bad(z1, y1) | z1;
    const uint32_t bad = ((c0 | c1) & NIB3) | ((y0 ^ 0x44444444u) & 0xEEEEEEEEu) | ((y1 ^ 0x44444444u) & 0xEEEEEEEEu);
    uint32_t vv = 0xFFFFu, e = CI ? 0xFFFFu : 0u;
    if (bad) { /* rare: N, lower case, bytes outside the arena */
        /* per nibble at bit 3: H = 01x? (f), upper case = f and H1 = 0 */
        const uint32_t f0 = ~y0 & (y0 << 1), f1 = ~y1 & (y1 << 1);
        const uint32_t vci0 = ~c0 & f0 & NIB3, vci1 = ~c1 & f1 & NIB3;
        const uint32_t vcs0 = vci0 & ~(y0 << 2), vcs1 = vci1 & ~(y1 << 2);
        if (CI) {
            vv = __byte_perm(vci0 * GATHER3, vci1 * GATHER3, 0x7373u);
            e = __byte_perm(vcs0 * GATHER3, vcs1 * GATHER3, 0x7373u);
        } else {
            /* 'N' = 0x4E: H = 0100, L = 1110 */
            const uint32_t n0 = f0 & ~(y0 << 2) & ~(y0 << 3) & z0 & (z0 << 1) & (z0 << 2) & ~(z0 << 3) & NIB3;
            const uint32_t n1 = f1 & ~(y1 << 2) & ~(y1 << 3) & z1 & (z1 << 1) & (z1 << 2) & ~(z1 << 3) & NIB3;
            vv = __byte_perm(vcs0 * GATHER3, vcs1 * GATHER3, 0x7373u);
            e = __byte_perm(n0 * GATHER3, n1 * GATHER3, 0x7373u);
        }
        l &= vv; h &= vv;
    }
    *lo = l; *hi = h; *v = vv; *ex = e;
}
__device__ __forceinline__ uint4 fill16() { return make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u); }

}  // namespace swar
