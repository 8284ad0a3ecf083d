"""SURVEY 8(a) row M / 8(f) #4b — the Matcher pass of FusionMapper::remove_alignables (src/core/fusion_mapper.rs:488-542,
src/core/matcher.rs).

CPU part: the literal oracle restatement (oracle/gf_oracle_matcher.cpp) against (a) hand-derived cases and (b) a second,
independent closed form written here from the analysis in include/genefuse_gpu.h — "parity unpinned": the reference holds
no test for this code and cannot be built here.
GPU part (-m gpu): gf_reference_create / gf_alignable_filter against the oracle through the C ABI.
"""
import os
import random

import numpy as np
import pytest

import _oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CODE = {ord("A"): 0, ord("T"): 1, ord("C"): 2, ord("G"): 3}
COMP = {ord("A"): "T", ord("a"): "T", ord("T"): "A", ord("t"): "A", ord("C"): "G", ord("c"): "G", ord("G"): "C", ord("g"): "C"}


# ---------------------------------------------------------------------------------- independent closed form (Python)
def kept_positions(contig):
    """positions index_contig_bytes keeps, with their key (matcher.rs:227-289 in closed form): i < len - 16, base i is ACGT
    (after upper-casing) and the rolling 32-bit value of the ACGT run up to i is < 4, i.e. the (at most 15) bases before i
    inside the run are all 'A'."""
    s = contig.upper()
    out = []
    run_start = None
    for i in range(max(0, len(s) - 16)):
        c = s[i]
        if c not in CODE:
            run_start = None
            continue
        if run_start is None:
            run_start = i
        lo = max(run_start, i - 15)
        if all(s[j] == ord("A") for j in range(lo, i)):
            out.append((i, CODE[c]))
    return out


def present_sets(seq):
    """codes at the k-mer starts [0, len-16] of the read (upper case only) and of its reverse complement"""
    f, r = set(), set()
    n = len(seq)
    for j in range(0, n - 15):
        if seq[j] in CODE:
            f.add(CODE[seq[j]])
    rc = "".join(COMP.get(b, "N") for b in reversed(seq)).encode()
    for j in range(0, n - 15):
        if rc[j] in CODE:
            r.add(CODE[rc[j]])
    return f, r


def model(contigs, seqs):
    """(key_positions, n_removed, panic_seq, bloom_bits, panic_stage) by the closed form"""
    for j, s in enumerate(seqs):
        if len(s) < 15:
            return None, 0, j, None, 4
    bloom = set()
    sets = [present_sets(s) for s in seqs]
    for f, r in sets:
        bloom |= f | r
    if any(len(c) < 16 for c in contigs):
        return None, 0, -1, sum(1 << k for k in bloom), 1
    lists = {k: [] for k in range(4)}
    for ci, c in enumerate(contigs):
        for pos, key in kept_positions(c):
            if key in bloom:
                lists[key].append((ci, pos))
    counts = tuple(len(lists[k]) for k in range(4))
    voting = {k for k in range(4) if 1 <= counts[k] <= 50 and any(not (c == 0 and p == j) for j, (c, p) in enumerate(lists[k]))}
    absent = {k for k in range(4) if counts[k] == 0}
    for j, (f, r) in enumerate(sets):
        if f & voting and f & absent:
            return counts, 0, j, sum(1 << k for k in bloom), 2
        if r & voting and r & absent:
            return counts, 0, j, sum(1 << k for k in bloom), 3
    return counts, 0, -1, sum(1 << k for k in bloom), 0


def check_against_model(contigs, seqs, got):
    counts, n_removed, panic_seq, bloom, stage = model(contigs, seqs)
    g_counts, g_removed, g_seq, g_bloom, g_stage = got
    assert g_stage == stage, (got, stage)
    assert g_removed == n_removed == 0
    assert g_seq == panic_seq
    if stage not in (1, 4):     # the index is complete
        assert g_counts == counts
    if stage != 4:
        assert g_bloom == bloom


def random_contig(rng, n, p_n=0.01, p_lower=0.1, a_runs=True):
    out = bytearray(rng.choice(b"ACGT") for _ in range(n))
    for i in range(n):
        x = rng.random()
        if x < p_n:
            out[i] = rng.choice(b"NRY-")
        elif x < p_n + p_lower:
            out[i] = out[i] | 0x20
    if a_runs and n > 200:
        for _ in range(max(1, n // 400)):
            a = rng.randrange(0, n - 40)
            ln = rng.randrange(10, 40)
            out[a:a + ln] = (b"A" if rng.random() < 0.7 else b"a") * ln
    return bytes(out)


def random_cases(seed, n_cases):
    rng = random.Random(seed)
    for _ in range(n_cases):
        contigs = [random_contig(rng, rng.choice([16, 17, 40, 300, 2000, 5000])) for _ in range(rng.randrange(1, 5))]
        if rng.random() < 0.1:
            contigs.insert(rng.randrange(len(contigs) + 1), b"ACGT" * rng.randrange(0, 4))     # shorter than 16
        seqs = []
        for _ in range(rng.randrange(0, 6)):
            ln = rng.choice([15, 16, 30, 100, 151, 270])
            alpha = rng.choice([b"ACGT", b"ACGTN", b"AC", b"acgtACGT", b"A", b"GT"])
            seqs.append(bytes(rng.choice(alpha) for _ in range(ln)))
        if rng.random() < 0.05:
            seqs.append(b"ACGT" * rng.randrange(0, 4))
        yield contigs, seqs


# ---------------------------------------------------------------------------------- CPU: oracle
def test_oracle_matcher_hand_cases():
    # one contig "AAAAAT" + 30 x C: rolling value < 4 at positions 0..5 (A A A A A T), then never again (the T stays in
    # the window for 15 positions, and the C run never restarts) -> key A: 5 positions, key T: 1, when both are in the bloom
    contig = b"AAAAAT" + b"C" * 30
    seqs = [b"A" * 16 + b"T" * 4]          # forward starts: A, then positions 1..4 -> A (len-16 = 4); rc = AAAA TTTT.. -> A, T? see model
    flags, res, rc = _oracle.remove_alignables([contig], seqs)
    check_against_model([contig], seqs, res.astuple())
    assert res.astuple()[0][0] == 5 and list(flags) == [0]
    # key A's five positions are (0,0),(0,1),...,(0,4): every vote packs to 0 (contig 0, position == list index) -> no votes
    # -> topcount[0] == 0 -> None, although key T (1 position: (0,5)) ... votes only if the read holds a T start
    assert res.panic_stage in (0, 2, 3)
    # a read that holds a voting key (T, one position, packs to 5 != 0) and an absent key (G) -> the reference panics
    seqs = [b"TTTTGGGG" + b"A" * 20]
    flags, res, rc = _oracle.remove_alignables([contig], seqs)
    check_against_model([contig], seqs, res.astuple())
    assert res.panic_stage == 2 and res.panic_seq == 0 and rc == -5
    # the same read, but the reference also has a G start -> nothing absent -> None for every read
    contig2 = contig + b"N" + b"G" + b"C" * 20
    flags, res, rc = _oracle.remove_alignables([contig2], seqs)
    check_against_model([contig2], seqs, res.astuple())
    assert res.panic_stage == 0 and rc == 0 and list(flags) == [0]


def test_oracle_matcher_tinyref_and_testdata():
    from genefuserust_b200.host import FastaReader, FastqReaderPair
    ref = FastaReader(os.path.join(HERE, "golden", "testdata", "tinyref.fa")).read_all()
    contigs = [ref.m_all_contigs[k] for k in sorted(ref.m_all_contigs)]
    (_n1, _n2), batch = FastqReaderPair(os.path.join(HERE, "golden", "testdata", "R1.fq"),
                                        os.path.join(HERE, "golden", "testdata", "R2.fq")).read_all()
    seqs = [batch.read(i, 1)[0] for i in range(batch.n)] + [batch.read(i, 2)[0] for i in range(batch.n)]
    flags, res, rc = _oracle.remove_alignables(contigs, seqs)
    check_against_model(contigs, seqs, res.astuple())
    # config 1 itself: no match survives, the Matcher sees no sequence, nothing is kept and nothing panics
    flags, res, rc = _oracle.remove_alignables(contigs, [])
    assert res.astuple() == ((0, 0, 0, 0), 0, -1, 0, 0) and rc == 0


def test_oracle_matcher_equals_closed_form_random():
    n = 0
    stages = set()
    for contigs, seqs in random_cases(20240201, 400):
        flags, res, rc = _oracle.remove_alignables(contigs, seqs)
        assert rc in (0, -5), rc          # -100 would mean a key >= 4 exists: the "degenerate" analysis itself is wrong
        check_against_model(contigs, seqs, res.astuple())
        assert not flags.any()
        stages.add(res.panic_stage)
        n += 1
    assert stages >= {0, 1, 2, 4}, stages


# ---------------------------------------------------------------------------------- GPU
def _gpu_matcher(contigs):
    from genefuserust_b200.host import Matcher
    return Matcher(contigs)


@pytest.mark.gpu
def test_alignable_filter_parity_random():
    stages = set()
    for k, (contigs, seqs) in enumerate(random_cases(777, 250)):
        m = _gpu_matcher(contigs)
        flags, res, rc = m.remove_alignables(seqs)
        oflags, ores, orc = _oracle.remove_alignables(contigs, seqs)
        assert rc == orc, (k, rc, orc)
        g, o = res.astuple(), ores.astuple()
        assert g[4] == o[4] and g[2] == o[2] and g[1] == o[1] == 0, (k, g, o)
        if g[4] not in (1, 4):
            assert g == o, (k, g, o)
        assert list(flags) == list(oflags)
        inf = m.info()
        assert inf.n_bases == sum(len(c) for c in contigs)
        stages.add(g[4])
        m.close()
    assert stages >= {0, 1, 2, 4}, stages


@pytest.mark.gpu
def test_alignable_filter_hand_cases_and_testdata():
    from genefuserust_b200.host import FastaReader, FastqReaderPair
    contig = b"AAAAAT" + b"C" * 30
    for contigs, seqs in (([contig], [b"TTTTGGGG" + b"A" * 20]),
                          ([contig + b"N" + b"G" + b"C" * 20], [b"TTTTGGGG" + b"A" * 20]),
                          ([contig], [b"A" * 16 + b"T" * 4]),
                          ([contig], [])):
        m = _gpu_matcher(contigs)
        flags, res, rc = m.remove_alignables(seqs)
        oflags, ores, orc = _oracle.remove_alignables(contigs, seqs)
        assert (rc, res.astuple()) == (orc, ores.astuple())
        m.close()
    ref = FastaReader(os.path.join(HERE, "golden", "testdata", "tinyref.fa")).read_all()
    m = _gpu_matcher(ref.m_all_contigs)
    (_n1, _n2), batch = FastqReaderPair(os.path.join(HERE, "golden", "testdata", "R1.fq"),
                                        os.path.join(HERE, "golden", "testdata", "R2.fq")).read_all()
    seqs = [batch.read(i, 1)[0] for i in range(batch.n)] + [batch.read(i, 2)[0] for i in range(batch.n)]
    contigs = [ref.m_all_contigs[k] for k in sorted(ref.m_all_contigs)]
    for ss in (seqs, []):
        flags, res, rc = m.remove_alignables(ss)
        oflags, ores, orc = _oracle.remove_alignables(contigs, ss)
        assert (rc, res.astuple()) == (orc, ores.astuple())
    m.close()


@pytest.mark.gpu
def test_alignable_filter_large_reference_host_and_device():
    """a 100 Mbase synthetic reference (several staging buffers, contigs cut across buffers, soft-masked stretches, N runs,
    poly-A): counts equal the oracle's; the same contigs resident in device memory give the same answer without a copy"""
    import torch
    rng = np.random.default_rng(5)
    lens = [61_000_003, 25_000_000, 13_999_981, 17, 16]
    contigs = []
    for ln in lens:
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, ln)]
        a = a.copy()
        # soft-masked (lower-case) stretches, N runs, poly-A
        for _ in range(max(1, ln // 2_000_000)):
            s = int(rng.integers(0, max(1, ln - 5000)))
            a[s:s + 3000] |= 0x20
            s = int(rng.integers(0, max(1, ln - 5000)))
            a[s:s + 500] = ord("N")
            s = int(rng.integers(0, max(1, ln - 5000)))
            a[s:s + 60] = ord("A")
        contigs.append(a)
    seqs = [b"ACGTTGCA" * 20, b"acgtNNNN" * 10 + b"ACGT" * 10]
    m = _gpu_matcher(contigs)
    flags, res, rc = m.remove_alignables(seqs)
    inf = m.info()
    oflags, ores, orc = _oracle.remove_alignables([c.tobytes() for c in contigs], seqs)
    assert (rc, res.astuple()) == (orc, ores.astuple())
    assert inf.h2d_bytes >= sum(lens) and inf.n_bases == sum(lens) and inf.short_contigs == 0
    assert sum(res.key_positions) > 100               # the N runs and poly-A stretches start ACGT runs / re-open the window
    m.close()
    dev = [torch.from_numpy(c).cuda() for c in contigs]
    torch.cuda.synchronize()
    m2 = _gpu_matcher([(t.data_ptr(), t.numel()) for t in dev])
    flags2, res2, rc2 = m2.remove_alignables(seqs)
    assert (rc2, res2.astuple()) == (rc, res.astuple())
    assert m2.info().h2d_bytes == 0
    m2.close()
