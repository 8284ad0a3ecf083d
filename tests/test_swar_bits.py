"""CPU model of the 16-base SWAR step of the read converters (csrc/gf_swar.cuh: block16): the multiply gathers, the
nibble-domain validity test and the slow path, checked against the plain definition of the planes
(make_kmer_bytes' code A0 T1 C2 G3, src/core/indexer.rs:888-904; reverse_complement's case rule, src/core/sequence.rs:52-60).
The constants are read from the CUDA source so that the model cannot drift from the kernel."""
import os
import random
import re

SRC = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "genefuserust_b200", "csrc", "gf_swar.cuh")).read()
M32 = 0xFFFFFFFF


def _const(pattern):
    m = re.search(pattern, SRC)
    assert m, pattern
    return [int(x, 16) for x in m.groups()]


MUL_LO, = _const(r"\(z0 & 0x44444444u\) \* (0x[0-9A-Fa-f]+)u")
MUL_HI, = _const(r"\(z0 & 0x22222222u\) \* (0x[0-9A-Fa-f]+)u")
NIB3, = _const(r"constexpr uint32_t NIB3 = (0x[0-9A-Fa-f]+)u;")
GATHER3, = _const(r"constexpr uint32_t GATHER3 = (0x[0-9A-Fa-f]+)u;")
FMT_XOR, FMT_AND = _const(r"\(\(y0 \^ (0x[0-9A-Fa-f]+)u\) & (0x[0-9A-Fa-f]+)u\)")


def byte_perm(x, y, s):
    b = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    return sum(b[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def bitsel(a, b, m):
    return (a & m) | (b & ~m & M32)


def nib_code_bad(z, y):
    l2, l1, l0, h0 = (z << 1) & M32, (z << 2) & M32, (z << 3) & M32, (y << 3) & M32
    return ((h0 ^ (l2 & ~l1)) | (~(l0 ^ h0) & M32)) & M32


def block16(x, ci):
    z0, z1 = bitsel(x[0], (x[1] << 4) & M32, 0x0F0F0F0F), bitsel(x[2], (x[3] << 4) & M32, 0x0F0F0F0F)
    y0, y1 = bitsel(x[0] >> 4, x[1], 0x0F0F0F0F), bitsel(x[2] >> 4, x[3], 0x0F0F0F0F)
    lo = byte_perm(((z0 & 0x44444444) * MUL_LO) & M32, ((z1 & 0x44444444) * MUL_LO) & M32, 0x7373)
    hi = byte_perm(((z0 & 0x22222222) * MUL_HI) & M32, ((z1 & 0x22222222) * MUL_HI) & M32, 0x7373)
    c0, c1 = nib_code_bad(z0, y0) | z0, nib_code_bad(z1, y1) | z1
    bad = ((c0 | c1) & NIB3) | ((y0 ^ FMT_XOR) & FMT_AND) | ((y1 ^ FMT_XOR) & FMT_AND)
    v, ex = 0xFFFF, (0xFFFF if ci else 0)
    if bad:
        f0, f1 = ~y0 & (y0 << 1) & M32, ~y1 & (y1 << 1) & M32
        vci0, vci1 = ~c0 & f0 & NIB3, ~c1 & f1 & NIB3
        vcs0, vcs1 = vci0 & ~(y0 << 2) & M32, vci1 & ~(y1 << 2) & M32
        g = lambda a, b: byte_perm((a * GATHER3) & M32, (b * GATHER3) & M32, 0x7373)
        if ci:
            v, ex = g(vci0, vci1), g(vcs0, vcs1)
        else:
            n = lambda z, y, f: f & ~(y << 2) & ~(y << 3) & z & (z << 1) & (z << 2) & ~(z << 3) & NIB3 & M32
            v, ex = g(vcs0, vcs1), g(n(z0, y0, f0), n(z1, y1, f1))
        lo &= v
        hi &= v
    return lo & 0xFFFF, hi & 0xFFFF, v & 0xFFFF, ex & 0xFFFF


def planes(bs, ci):
    lo = hi = v = ex = 0
    for p, b in enumerate(bs):
        ch = chr(b)
        up = ch.upper() if ci else ch
        if up in "ACGT":
            code = "ATCG".index(up)
            lo |= (code & 1) << p
            hi |= (code >> 1) << p
            v |= 1 << p
            if ci and ch in "ACGT":
                ex |= 1 << p
        if not ci and ch == "N":
            ex |= 1 << p
    return lo, hi, v, ex


def _words(bs):
    return [int.from_bytes(bs[4 * j:4 * j + 4], "little") for j in range(4)]


def test_block16_random_blocks():
    rng = random.Random(1)
    alpha = b"ACGTACGTACGTACGTNacgtnRY\x00\xff@BDEFPQSUVW"
    for it in range(40000):
        mode = it % 4
        if mode == 0:
            bs = bytes(rng.choice(b"ACGT") for _ in range(16))
        elif mode == 1:
            bs = bytes(rng.choice(alpha) for _ in range(16))
        elif mode == 2:
            bs = bytes(rng.randrange(256) for _ in range(16))
        else:
            bs = bytes(rng.choice(b"ACGTacgt") for _ in range(16))
        for ci in (False, True):
            assert block16(_words(bs), ci) == planes(bs, ci), (bs, ci)


def test_block16_every_byte_value_in_every_position():
    for pos in range(16):
        for b in range(256):
            bs = bytearray(b"ACGTTGCAACGTTGCA")
            bs[pos] = b
            for ci in (False, True):
                assert block16(_words(bytes(bs)), ci) == planes(bytes(bs), ci), (pos, b, ci)


# ---- the reference scan's converter (csrc/gf_matcher.cu: group_va / sector_va): valid (ACGT in either case) and isA planes of
# one 32-byte sector, branch-free in the nibble domain
MSRC = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "genefuserust_b200", "csrc", "gf_matcher.cu")).read()


def _mconst(pattern):
    m = re.search(pattern, MSRC)
    assert m, pattern
    return [int(x, 16) for x in m.groups()]


VA_MASK, = _mconst(r"t3 = t1 & t2 & (0x[0-9A-Fa-f]+)u;")
VA_MUL, = _mconst(r"constexpr uint32_t M = (0x[0-9A-Fa-f]+)u;")
VA_IN, VA_OUT = _mconst(r"\*v = __byte_perm\(__byte_perm\(v0 \* M, v1 \* M, (0x[0-9A-Fa-f]+)u\), __byte_perm\(v2 \* M, v3 \* M, 0x[0-9A-Fa-f]+u\), (0x[0-9A-Fa-f]+)u\);")


def group_va(w0, w1):
    z = (w0 & 0x0F0F0F0F) | ((w1 << 4) & 0xF0F0F0F0)
    y = ((w0 >> 4) & 0x0F0F0F0F) | (w1 & 0xF0F0F0F0)
    l2, l1, l0, h2, h0 = (z << 1) & M32, (z << 2) & M32, (z << 3) & M32, (y << 1) & M32, (y << 3) & M32
    t1 = ~y & h2 & ~z & M32
    t2 = ~(h0 ^ (l2 & ~l1)) & M32
    vv = t1 & t2 & VA_MASK & (l0 ^ h0)
    return vv, vv & ~l2 & ~l1 & M32


def sector_va(bs):
    w = [int.from_bytes(bs[4 * j:4 * j + 4], "little") for j in range(8)]
    g = [group_va(w[2 * k], w[2 * k + 1]) for k in range(4)]
    out = []
    for sel in (0, 1):
        m = [(g[k][sel] * VA_MUL) & M32 for k in range(4)]
        out.append(byte_perm(byte_perm(m[0], m[1], VA_IN), byte_perm(m[2], m[3], VA_IN), VA_OUT))
    return tuple(out)


def va_planes(bs):
    v = pa = 0
    for p, b in enumerate(bs):
        up = chr(b).upper() if b < 128 else "?"
        if up in "ACGT":
            v |= 1 << p
        if up == "A":
            pa |= 1 << p
    return v, pa


def test_sector_va_every_byte_value_in_every_position():
    for pos in range(32):
        for b in range(256):
            bs = bytearray(b"ACGTTGCAacgtTGCAAAAACCCCGGGGTTTT")
            bs[pos] = b
            assert sector_va(bytes(bs)) == va_planes(bytes(bs)), (pos, b)


def test_sector_va_random_sectors():
    rng = random.Random(2)
    alpha = b"ACGTACGTAAAAacgtaNnRY\x00\xff@BDEFPQSUVWdeqsuvw"
    for it in range(20000):
        if it % 3 == 0:
            bs = bytes(rng.choice(b"ACGTacgt") for _ in range(32))
        elif it % 3 == 1:
            bs = bytes(rng.choice(alpha) for _ in range(32))
        else:
            bs = bytes(rng.randrange(256) for _ in range(32))
        assert sector_va(bs) == va_planes(bs), bs
