"""The C++ oracle against a second, independently written Python restatement of the same reference code
(tests/ref_port.py) on randomised fusion reads — including repeats, N, lower case, reversed genes and noisy reads."""
import random

import _oracle as orc
import ref_port
from genefuserust_b200._abi import gf_params

COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def _genes(rng, n=5):
    rnd = lambda k: bytes(rng.choice(b"ACGT") for _ in range(k))
    genes = []
    shared = rnd(120)                                    # NORMAL dupes across genes
    many = rnd(60)                                       # HIGH dupes (7 copies)
    for g in range(n):
        s = bytearray(rnd(rng.randint(400, 900)))
        if g < 3:
            s[100:220] = shared
        for k in range(2 if g < 4 else 0):
            p = 250 + 70 * k
            s[p:p + 60] = many
        if g == 1:
            s[300:330] = b"N" * 30
        genes.append((bytes(s), bool(g % 2)))
    return genes


def _reads(rng, genes, n):
    out = []
    for _ in range(n):
        (a, _), (b, _) = rng.choice(genes), rng.choice(genes)
        x, y = rng.randint(30, 110), rng.randint(30, 110)
        pa, pb = rng.randrange(0, len(a) - x), rng.randrange(0, len(b) - y)
        left, right = a[pa:pa + x], b[pb:pb + y]
        if rng.random() < 0.5:
            left = left[::-1].translate(COMP)
        if rng.random() < 0.5:
            right = right[::-1].translate(COMP)
        frag = bytearray(left + right)
        for _ in range(rng.choice((0, 0, 0, 1, 2, 6))):
            frag[rng.randrange(len(frag))] = rng.choice(b"ACGTNacgt")
        if rng.random() < 0.15:
            p = rng.randrange(len(frag))
            frag[p:p] = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 3)))   # insertion
        if rng.random() < 0.5:
            frag = bytearray(bytes(frag[::-1]).translate(COMP))
        out.append(bytes(frag))
    return out


def test_cpp_oracle_equals_python_port():
    rng = random.Random(20240201)
    n_match = n_mapable_only = 0
    for rep in range(6):
        genes = _genes(rng)
        ref = ref_port.RefIndexer(genes)
        idx = orc.OracleIndex(genes)
        c = idx.counts()
        kinds = [v[0] for v in ref.kmer_pos.values()]
        assert c["n_keys"] == len(ref.kmer_pos)
        assert c["n_high"] == sum(1 for k in kinds if k == -2) and c["n_normal"] == sum(1 for k in kinds if k == -1)
        for seq in _reads(rng, genes, 250):
            want_segs = [(s, e, gp[0], gp[1]) for (s, e, gp) in ref.map_read(seq)]
            assert idx.map_read(seq) == want_segs, seq
            want, want_mapable = ref.fusion_map_read(seq)
            m, mapable = idx.fusion_map_read(seq)
            assert mapable == want_mapable
            got = None if m is None else (m.read_break, m.l_contig, m.l_pos, m.r_contig, m.r_pos, m.gap, m.l_dist,
                                          m.r_dist, m.seq_len)
            assert got == want, (seq, got, want)
            n_match += want is not None
            n_mapable_only += want is None and want_mapable
        idx.close()
    assert n_match > 100 and n_mapable_only > 30, (n_match, n_mapable_only)


def test_thresholds_in_both_restatements():
    rng = random.Random(5)
    genes = _genes(rng)
    for thr in (1, 2, 5, 7):
        p = gf_params.default()
        p.skip_key_dup_threshold = thr
        ref = ref_port.RefIndexer(genes, dup_threshold=thr)
        idx = orc.OracleIndex(genes, params=p)
        c = idx.counts()
        kinds = [v[0] for v in ref.kmer_pos.values()]
        assert (c["n_keys"], c["n_high"], c["n_normal"]) == (len(kinds), kinds.count(-2), kinds.count(-1)), thr
        idx.close()


def test_scan_pair_end_policy_in_both_restatements():
    """merge -> map -> rc retry -> reversed flag (pescanner.rs:427-518): the C++ oracle's batch scan against the Python port"""
    from genefuserust_b200 import ReadBatch
    rng = random.Random(77)
    genes = _genes(rng, n=6)
    ref = ref_port.RefIndexer(genes)
    idx = orc.OracleIndex(genes)
    r1s, r2s = [], []
    for frag in _reads(rng, genes, 400):
        L = rng.choice((75, 100, 150))
        s1 = frag[:L]
        s2 = frag[::-1].translate(COMP)[:L]
        q = lambda n: bytes(rng.choice(b"EEEEEA/?0") for _ in range(n))
        r1s.append((s1, q(len(s1))))
        r2s.append((s2, q(len(s2))))
    b = ReadBatch.from_reads(r1s, r2s)
    got = idx.scan(b, threads=2)
    want = []
    for i, ((s1, q1), (s2, q2)) in enumerate(zip(r1s, r2s)):
        for rec in ref_port.scan_pair(ref, s1, q1, s2, q2):
            want.append((i,) + rec)
    # oracle tuple: pair, source, used_rc, reversed, read_break, lc, lp, rc, rp, gap, ld, rd, seq_len, olen, diff, flags
    assert [g[:13] for g in got] == want
    assert len(want) > 60 and any(w[2] for w in want) and any(w[1] == 0 for w in want) and any(w[1] == 2 for w in want)
    idx.close()


def test_adjust_fusion_break_in_both_restatements():
    """FusionResult::adjust_fusion_break / get_ref_seq (fusion_result.rs:299-397, 770-798): C++ oracle vs the independent
    Python port on randomised junction reads (substitutions, indels near the break, wrong breaks, short / empty references)"""
    import random
    rng = random.Random(17)
    rnd = lambda n: bytes(rng.choice(b"ACGT") for _ in range(n))
    n_shift = 0
    for it in range(400):
        gl, gr = rnd(400), rnd(400)
        ll, rl = rng.randint(25, 140), rng.randint(25, 140)
        a, b = rng.randint(0, 400 - ll), rng.randint(0, 400 - rl)
        read = bytearray(gl[a:a + ll] + gr[b:b + rl])
        for _ in range(rng.choice((0, 0, 1, 2, 3))):
            p = rng.randrange(len(read))
            r = rng.random()
            if r < 0.6:
                read[p] = rng.choice(b"ACGTN")
            elif r < 0.8:
                del read[p]
            else:
                read.insert(p, rng.choice(b"ACGT"))
        read = bytes(read)
        rb = max(0, min(len(read) - 1, ll - 1 + rng.choice((0, 0, 0, -1, 1, -2, 2, -3, 3, -5, 5))))
        lref = gl[max(0, a - rng.randint(0, 30)):a + ll] if it % 7 else b""
        rref = gr[b:b + rl + rng.randint(0, 30)] if it % 11 else gr[b:b + 5]
        want = ref_port.adjust_fusion_break(read, rb, lref, rref)
        got = orc.adjust_fusion_break(read, rb, lref, rref)
        assert got == ((0, 0, 0, 1) if want is None else want + (0,)), (it, rb, len(read), got, want)
        n_shift += want is not None and want[0] != 0
        # get_ref_seq on both strands and at the edges
        s0 = rng.randint(-420, 420)
        e0 = s0 + rng.randint(-5, 60)
        assert orc.get_ref_seq(gl, s0, e0) == ref_port.get_ref_seq(gl, s0, e0) or e0 < s0, (s0, e0)
    assert n_shift > 50
