"""The device inflate routine (csrc/gf_inflate.cuh: one GPU thread per BGZF member) compiled for the host and checked against
zlib: raw DEFLATE streams of FASTQ-like text, random bytes, runs and empty input at levels 0 / 1 / 6 / 9 with the default,
fixed-Huffman, Huffman-only and RLE strategies (stored, fixed and dynamic blocks), the CRC-32 routine, and that a member whose
sizes do not come out exactly is reported (flate2's MultiGzDecoder behind FastqReader fails on those too,
src/core/fastq_reader.rs:39-69)."""
import ctypes as C
import os
import random
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "inflate_host.cpp")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = os.path.join(str(tmp_path_factory.mktemp("inflate")), "libgf_inflate_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wall", SRC, "-o", so])
    lib = C.CDLL(so)
    lib.gf_test_inflate.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]
    lib.gf_test_inflate.restype = C.c_int
    lib.gf_test_crc32.argtypes = [C.c_char_p, C.c_uint32]
    lib.gf_test_crc32.restype = C.c_uint32
    lib.gf_test_crc32_sliced.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
    lib.gf_test_crc32_sliced.restype = C.c_uint32
    return lib


def fastq(rng, n):
    out = []
    for i in range(n):
        L = rng.choice((36, 75, 100, 150, 151, 250))
        out.append(b"@read%d/1\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(L)),
                                               bytes(rng.choice(b"EEEEEEA/<6") for _ in range(L))))
    return b"".join(out)


def raw_deflate(data, level, strategy):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def test_inflate_equals_zlib(lib):
    rng = random.Random(1)
    cases = 0
    for trial in range(36):
        kind = trial % 6
        if kind == 0:
            data = fastq(rng, rng.randint(1, 300))[:65280]
        elif kind == 1:
            data = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 70000)))[:65536]
        elif kind == 2:
            data = b"A" * rng.randint(1, 65536)
        elif kind == 3:
            data = (b"ACGT" * 20000)[:rng.randint(1, 65536)]
        elif kind == 4:
            data = b""
        else:
            data = fastq(rng, 200)[:rng.randint(1, 60000)]
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                comp = raw_deflate(data, level, strat)
                out = C.create_string_buffer(max(1, len(data)))
                assert lib.gf_test_inflate(comp, len(comp), out, len(data)) == 0, (trial, level, strat)
                assert out.raw[:len(data)] == data, (trial, level, strat)
                assert lib.gf_test_crc32(data, len(data)) == zlib.crc32(data)
                cases += 1
                if len(data) > 10:   # sizes that do not come out exactly are errors, never silently accepted
                    assert lib.gf_test_inflate(comp, len(comp), out, len(data) - 1) != 0
                    assert lib.gf_test_inflate(comp, len(comp) - 1, out, len(data)) != 0
                    assert lib.gf_test_inflate(comp + b"\0", len(comp) + 1, out, len(data)) != 0
    assert cases == 36 * 16


def test_inflate_rejects_garbage(lib):
    rng = random.Random(2)
    data = fastq(rng, 100)
    comp = bytearray(raw_deflate(data, 6, zlib.Z_DEFAULT_STRATEGY))
    out = C.create_string_buffer(len(data))
    bad = 0
    for _ in range(300):
        c = bytearray(comp)
        for _ in range(rng.randint(1, 4)):
            c[rng.randrange(len(c))] ^= 1 << rng.randrange(8)
        rc = lib.gf_test_inflate(bytes(c), len(c), out, len(data))
        # a flipped bit either breaks the stream (reported) or yields other text of the same length (the CRC catches that)
        if rc != 0 or out.raw != data:
            bad += 1
    assert bad >= 295
    for _ in range(100):   # random bytes: must return (no crash, no endless loop)
        junk = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 400)))
        lib.gf_test_inflate(junk, len(junk), out, len(data))


def test_crc32_from_slices(lib):
    """the device computes a member's CRC-32 as 32 lanes x one slice each and combines them (crc_multmodp / crc_x8n, zlib's
    crc32_combine arithmetic): any length, any number of slices"""
    rng = random.Random(3)
    for n in (0, 1, 2, 31, 32, 33, 100, 4096, 65280, 65536, 70001):
        data = bytes(rng.getrandbits(8) for _ in range(n))
        for parts in (1, 2, 3, 32):
            assert lib.gf_test_crc32_sliced(data, n, parts) == zlib.crc32(data), (n, parts)
