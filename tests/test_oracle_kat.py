"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the path
(SURVEY.md §8c).  Vectors were lifted by tests/golden/make_fixtures.py; each cites file:line."""
import json
import os
import random

import _oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLD, "kat.json")))


def read_fastq(path):
    lines = open(path, "rb").read().split(b"\n")
    recs = []
    for i in range(0, len(lines) - 3, 4):
        recs.append((lines[i], lines[i + 1], lines[i + 3]))
    return recs


def test_fast_merge_assert():
    # src/core/read.rs:450-486
    k = KAT["fast_merge"]
    r = orc.fast_merge(k["r1_seq"].encode(), k["r1_qual"].encode(), k["r2_seq"].encode(), k["r2_qual"].encode())
    assert r is not None
    assert r[0].decode() == k["merged_seq"]


def test_merged_fixture_testdata_pair1():
    # src/core/indexer.rs:1059 holds fast_merge(testdata pair #1) as a literal
    k = KAT["merged_fixture"]
    r1 = read_fastq(os.path.join(GOLD, "testdata", "R1.fq"))
    r2 = read_fastq(os.path.join(GOLD, "testdata", "R2.fq"))
    assert len(r1) == len(r2) == 3
    seq, qual, olen, diff = orc.fast_merge(r1[0][1], r1[0][2], r2[0][1], r2[0][2])
    assert seq.decode() == k["seq"]
    assert qual.decode() == k["qual"]
    assert (r1[0][0].decode() + f" merged_diff_{diff}") == k["name"]
    # pairs #2/#3 merge with overlap 138 -> 161 bp
    for i in (1, 2):
        seq, qual, olen, diff = orc.fast_merge(r1[i][1], r1[i][2], r2[i][1], r2[i][2])
        assert (olen, len(seq)) == (138, 161)


def test_edit_distance_vectors():
    # src/core/edit_distance.rs:221-261
    k = KAT["edit_distance"]
    for a, b, e in zip(k["a"], k["b"], k["expect"]):
        assert orc.edit_distance(a.encode(), b.encode()) == e
        assert orc.levenshtein_dp(a.encode(), b.encode()) == e


def test_edit_distance_bitvector_equals_dp():
    rng = random.Random(7)
    for _ in range(1500):
        n = rng.randint(0, 660)
        a = bytes(rng.choice(b"ACGTN") for _ in range(n))
        b = bytearray(a)
        for _ in range(rng.randint(0, 12)):
            if not b:
                break
            p = rng.randrange(len(b))
            op = rng.random()
            if op < 0.5:
                b[p] = rng.choice(b"ACGT")
            elif op < 0.75:
                del b[p]
            else:
                b.insert(p, rng.choice(b"ACGT"))
        if rng.random() < 0.1:
            b = bytearray(rng.choice(b"ACGT") for _ in range(rng.randint(0, 660)))
        got = orc.edit_distance(a, bytes(b))
        want = orc.levenshtein_dp(a, bytes(b))
        if min(len(a), len(b)) > 640:
            assert got == -1000000 - want  # the reference would panic here (edit_distance.rs:94-100,177-196)
        else:
            assert got == want, (len(a), len(b))


def test_reverse_complement():
    # src/core/sequence.rs:66-70
    for src, want in KAT["reverse_complement"]["pairs"]:
        assert orc.reverse_complement(src.encode()).decode() == want
    assert orc.reverse_complement(b"acgtnRyX") == b"NNNNACGT"


def test_gp_roundtrip():
    # src/core/indexer.rs:981-1016
    import ctypes as C
    L = orc.lib()
    k = KAT["gp_roundtrip"]
    for c, p in zip(k["contigs"], k["positions"]):
        v = L.orc_gp_to_i64(c, p)
        assert v == ((c << 32) | (p & 0xFFFFFFFF)) or c < 0
        c2, p2 = C.c_int16(), C.c_int32()
        L.orc_i64_to_gp(v, C.byref(c2), C.byref(p2))
        assert (c2.value, p2.value) == (c, p)
        assert L.orc_gp_to_i64(c2.value, p2.value) == v


def test_make_kmer_code():
    # src/core/indexer.rs:852-913: A=0 T=1 C=2 G=3, first base in the top bits; anything else -> -1
    L = orc.lib()
    assert L.orc_make_kmer(b"AAAAAAAAAAAAAAAA", 0) == 0
    assert L.orc_make_kmer(b"GGGGGGGGGGGGGGGG", 0) == 0xFFFFFFFF
    assert L.orc_make_kmer(b"TAAAAAAAAAAAAAAC", 0) == (1 << 30) | 2
    assert L.orc_make_kmer(b"AAAAAAAANAAAAAAA", 0) == -1
    assert L.orc_make_kmer(b"AAAAAAAAaAAAAAAA", 0) == -1
    assert L.orc_make_kmer(b"NACGTACGTACGTACGT", 1) == L.orc_make_kmer(b"ACGTACGTACGTACGT", 0)


def test_adjust_fusion_break_hand_cases():
    """FusionResult::adjust_fusion_break (fusion_result.rs:299-397) on cases small enough to verify by hand"""
    left_gene = b"ACGTTGCAAGGCTTAACCGGATCGATCGTTAGC"      # 33 bases
    right_gene = b"TTGACCATGGCAATCGGATTACAGGCTTACGAT"     # 33 bases
    # a read that is exactly left_gene[3:33] + right_gene[0:28]: break at 29, no shift needed, both distances 0
    read = left_gene[3:] + right_gene[:28]
    assert orc.adjust_fusion_break(read, 29, left_gene[3:], right_gene[:28]) == (0, 0, 0, 0)
    # the caller's break is 2 too far left: the best shift is +2 and the distances return to 0
    assert orc.adjust_fusion_break(read, 27, left_gene[3:], right_gene[:28]) == (2, 0, 0, 0)
    # ... 3 too far right: shift -3
    assert orc.adjust_fusion_break(read, 32, left_gene[3:], right_gene[:28]) == (-3, 0, 0, 0)
    # one substitution on the right side, 5 bases after the break: shift 0, right distance 1
    mut = bytearray(read)
    mut[35] = ord("A") if mut[35] != ord("A") else ord("C")
    assert orc.adjust_fusion_break(bytes(mut), 29, left_gene[3:], right_gene[:28]) == (0, 0, 1, 0)
    # empty references (get_ref_seq returned ""): every shift scores 0, the first one (-3) wins (strict '<' from 0xFFFF)
    assert orc.adjust_fusion_break(read, 29, b"", b"") == (-3, 0, 0, 0)
    # a shifted break outside the read is outside the reference's defined behaviour
    assert orc.adjust_fusion_break(read, 1, left_gene[3:], right_gene[:28])[3] == 1
    # get_ref_seq: forward slice, reverse strand = reverse complement of [-end, -start], straddling / overflow -> ""
    assert orc.get_ref_seq(left_gene, 3, 7) == left_gene[3:8]
    assert orc.get_ref_seq(left_gene, -7, -3) == orc.reverse_complement(left_gene[3:8])
    assert orc.get_ref_seq(left_gene, -2, 3) == b"" and orc.get_ref_seq(left_gene, 0, 5) == b""
    assert orc.get_ref_seq(left_gene, 30, 33) == b""
