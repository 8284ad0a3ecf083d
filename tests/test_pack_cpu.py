"""CPU test of the host packer of the packed upload (csrc/gf_pack.cpp, gf_pack_reads): the plane words of ragged reads with
N / lower case / IUPAC / arbitrary bytes against the plain definition of the planes (code A0 T1 C2 G3,
src/core/indexer.rs:888-904; R1 validity = upper-case ACGT + the 'N' plane fast_merge looks at, read.rs:339-367; R2 validity =
ACGT in either case + upper-case plane, sequence.rs:52-60), for both mates and several thread counts.  No device involved."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from genefuserust_b200._abi import load_library


def model(reads, mate2):
    words, woff, xwords, xoff = [], [], [], []
    for r in reads:
        nw = (len(r) + 31) // 32
        lo, hi, v, a = [0] * nw, [0] * nw, [0] * nw, [0] * nw
        flagged = False
        for p, b in enumerate(r):
            ch = chr(b) if b < 128 else "?"
            up = ch in "ACGT"
            val = (ch.upper() in "ACGT" and b < 128) if mate2 else up
            if not up:
                flagged = True
            k, bit = p >> 5, 1 << (p & 31)
            if val:
                code = "ATCG".index(ch.upper())
                lo[k] |= bit * (code & 1)
                hi[k] |= bit * (code >> 1)
                v[k] |= bit
            if (up if mate2 else ch == "N"):
                a[k] |= bit
        woff.append(len(words))
        words += lo + hi
        if flagged:
            xoff.append(len(xwords) + 1)
            xwords += v + a
        else:
            xoff.append(0)
    return words, woff, xwords, xoff


def run(lib, reads, mate2):
    seq = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if reads else np.zeros(0, np.uint8)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    off += np.uint64(1000)                      # offsets need not start at 0
    cap = 2 * sum((len(r) + 31) // 32 for r in reads)
    words = np.full(cap + 8, 0xDEADBEEF, dtype=np.uint32)
    xwords = np.full(cap + 8, 0xDEADBEEF, dtype=np.uint32)
    woff = np.zeros(len(reads), dtype=np.uint32)
    xoff = np.zeros(len(reads), dtype=np.uint32)
    nw, nx = C.c_uint64(0), C.c_uint64(0)
    rc = lib.gf_pack_reads(seq.ctypes.data, off.ctypes.data, len(reads), int(mate2), words.ctypes.data, woff.ctypes.data,
                           xwords.ctypes.data, xoff.ctypes.data, cap, C.byref(nw), C.byref(nx))
    assert rc == 0, lib.gf_last_error()
    assert nw.value == cap
    assert (words[cap:] == 0xDEADBEEF).all() and (xwords[cap:] == 0xDEADBEEF).all()    # nothing beyond the capacity is touched
    return list(map(int, words[:cap])), list(map(int, woff)), list(map(int, xwords[:nx.value])), list(map(int, xoff))


@pytest.mark.parametrize("threads", ["1", "4", "7"])
def test_pack_reads_against_the_plane_definition(threads, monkeypatch):
    lib = load_library()
    monkeypatch.setenv("GF_HOST_PACK", "1")        # (without it the packer is only offered with >= 8 packing threads)
    if not lib.gf_pack_supported():
        pytest.skip("no AVX-512BW on this host: the packed upload is not offered")
    monkeypatch.setenv("GF_PACK_THREADS", threads)
    rng = random.Random(int(threads))
    reads = []
    for k in range(3000):
        n = rng.choice((0, 1, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 150, 151, 200, 255, 256, 300, 1000)) if k % 3 else 150
        if k % 4 == 0:
            r = bytes(rng.choice(b"ACGT") for _ in range(n))
        elif k % 4 == 1:
            r = bytes(rng.choice(b"ACGTACGTACGTNacgtnRYK@[`{") for _ in range(n))
        elif k % 4 == 2:
            r = bytes(rng.randrange(256) for _ in range(n))
        else:
            r = bytearray(rng.choice(b"ACGT") for _ in range(n))
            if n:
                r[rng.randrange(n)] = rng.choice(b"Nn\x00\xffacgt")
            r = bytes(r)
        reads.append(r)
    for mate2 in (False, True):
        assert run(lib, reads, mate2) == model(reads, mate2), mate2
    assert run(lib, [], False) == ([], [], [], [])


def test_pack_reads_capacity_and_switch(monkeypatch):
    lib = load_library()
    monkeypatch.setenv("GF_HOST_PACK", "1")
    if not lib.gf_pack_supported():
        pytest.skip("no AVX-512BW on this host")
    seq = np.frombuffer(b"ACGT" * 40, dtype=np.uint8).copy()
    off = np.array([0, 160], dtype=np.uint64)
    buf = np.zeros(64, dtype=np.uint32)
    nw, nx = C.c_uint64(0), C.c_uint64(0)
    rc = lib.gf_pack_reads(seq.ctypes.data, off.ctypes.data, 1, 0, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data,
                           4, C.byref(nw), C.byref(nx))
    assert rc == -3 and nw.value == 10          # GF_E_CAPACITY, needed count
    monkeypatch.setenv("GF_HOST_PACK", "0")
    assert lib.gf_pack_supported() == 0
    monkeypatch.delenv("GF_HOST_PACK")
    monkeypatch.setenv("GF_PACK_THREADS", "3")     # many ranks on few cores: packing would cost more than it saves
    assert lib.gf_pack_supported() == 0
    monkeypatch.setenv("GF_PACK_THREADS", "12")
    assert lib.gf_pack_supported() == 1
