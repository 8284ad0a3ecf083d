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

/* The converters read one aligned 32-byte sector (two 16-byte blocks) per step and turn each block into 16 plane bits:
 *   - the code bits (bit 2 / bit 1 of the ASCII byte = A0 T1 C2 G3) of 8 bytes are gathered by ONE multiply:
 *     z = low nibbles of word 0 | low nibbles of word 1 << 4, (z & 0x44444444) * 0x00408102 has the eight bit-2 values in
 *     base order in its top byte (0x00810204 for bit 1); the partial products never collide, so there are no carries;
 *   - validity: the expected letter is looked up with PRMT from the low 3 bits of each byte (A 001, C 011, T 100, G 111)
 *     and xor-ed with the byte; a block whose 16 differences are all zero takes the fast path (valid = 0xFFFF);
 *   - misalignment a (0..31) of the read w.r.t. the sectors is removed in the bit domain: planes are built at bit position
 *     u = p + a and aligned word w = funnelshift(U[w], U[w+1], a). */
__device__ __forceinline__ uint32_t expect4(uint32_t x) {
    const uint32_t t = x & 0x07070707u;
    const uint32_t u = t | (t >> 4);
    return __byte_perm(0x43414141u, 0x47414154u, __byte_perm(u, 0u, 0x4420u));
}
/* One 16-byte block -> 16 plane bits in the LOW half of each result (the upper halves are garbage):
 *   lo / hi = code bits, v = valid (ACGT; either case when CI), ex = !CI: the byte is 'N';  CI: valid AND upper case */
template <bool CI>
__device__ __forceinline__ void block16(const uint4& x, uint32_t* lo, uint32_t* hi, uint32_t* v, uint32_t* ex) {
    const uint32_t z0 = (x.x & 0x0F0F0F0Fu) | ((x.y << 4) & 0xF0F0F0F0u);
    const uint32_t z1 = (x.z & 0x0F0F0F0Fu) | ((x.w << 4) & 0xF0F0F0F0u);
    uint32_t l = __byte_perm((z0 & 0x44444444u) * 0x00408102u, (z1 & 0x44444444u) * 0x00408102u, 0x7373u);
    uint32_t h = __byte_perm((z0 & 0x22222222u) * 0x00810204u, (z1 & 0x22222222u) * 0x00810204u, 0x7373u);
    constexpr uint32_t CM = CI ? 0xDFDFDFDFu : 0xFFFFFFFFu;
    const uint32_t d0 = (x.x ^ expect4(x.x)) & CM, d1 = (x.y ^ expect4(x.y)) & CM;
    const uint32_t d2 = (x.z ^ expect4(x.z)) & CM, d3 = (x.w ^ expect4(x.w)) & CM;
    uint32_t bad = d0 | d1 | d2 | d3;
    if (CI) bad |= (x.x | x.y | x.z | x.w) & 0x20202020u; /* a lower-case letter: ex differs from v */
    uint32_t vv = 0xFFFFu, e = CI ? 0xFFFFu : 0u;
    if (bad) { /* rare: N, lower case, bytes outside the arena */
        const uint32_t y0 = (zero_bytes(d0) >> 5) | (zero_bytes(d1) >> 1);
        const uint32_t y1 = (zero_bytes(d2) >> 5) | (zero_bytes(d3) >> 1);
        vv = __byte_perm(y0 * 0x00408102u, y1 * 0x00408102u, 0x7373u);
        if (CI) {
            const uint32_t w0 = ((x.x >> 3) & 0x04040404u) | ((x.y << 1) & 0x40404040u);
            const uint32_t w1 = ((x.z >> 3) & 0x04040404u) | ((x.w << 1) & 0x40404040u);
            e = vv & ~__byte_perm(w0 * 0x00408102u, w1 * 0x00408102u, 0x7373u);
        } else {
            const uint32_t n0 = (zero_bytes(x.x ^ 0x4E4E4E4Eu) >> 5) | (zero_bytes(x.y ^ 0x4E4E4E4Eu) >> 1);
            const uint32_t n1 = (zero_bytes(x.z ^ 0x4E4E4E4Eu) >> 5) | (zero_bytes(x.w ^ 0x4E4E4E4Eu) >> 1);
            e = __byte_perm(n0 * 0x00408102u, n1 * 0x00408102u, 0x7373u);
        }
        l &= vv; h &= vv;
    }
    *lo = l; *hi = h; *v = vv; *ex = e;
}
__device__ __forceinline__ uint4 fill16() { return make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u); }

}  // namespace swar
