"""CPU property test of the screen's drop rule (csrc/gf_screen_split.cuh), independent of the GPU:

    a k-mer votes at most once for any one diagonal, so with s_i = sites of the k-mer at even offset i,
    P = #{s_i >= 1}, T = sum min(s_i, 2), c_d = votes on diagonal d:
        count1 <= P,   count1 + count2 <= T,   count2 <= T - c_d  for every d.

For random on-target / fusion / repeat reads the votes of pass 1 (src/core/indexer.rs:277-346) are rebuilt from the
oracle's index lookups; whenever the gate (:353-360) would pass, every drop condition of k_diag (for ANY seed diagonal)
and of k_scan must be false — with exact site counts, i.e. the filter's false positives can only add to P and T.
"""
import random

import numpy as np

import _oracle as orc
from genefuserust_b200 import synth

NEED_MAJOR, NEED_MINOR = 20, 10          # ceil(40 / 2), ceil(20 / 2): major / minor_gene_key_requirement


def _panel():
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    genes = [synth.random_bases(7000 + g, 3000) for g in range(12)]
    a, b, c = (synth.random_bases(800 + k, n) for k, n in enumerate((900, 700, 500)))
    for g in (0, 1, 2):
        genes[g][500 + 100 * g:500 + 100 * g + len(a)] = a                    # 3 copies: NORMAL dupes
    genes[3][1000:1000 + len(b)] = b
    genes[4][200:200 + len(b)] = np.frombuffer(b.tobytes()[::-1].translate(comp), dtype=np.uint8)   # rc copy
    for g in (5, 6, 7, 8, 9, 10):
        genes[g][1500:1500 + len(c)] = c                                       # 6 copies: HIGH
    return [(g.tobytes(), bool(i % 2)) for i, g in enumerate(genes)]


def _votes(o, seq):
    """pass 1 of Indexer::map_read from index lookups: {(contig, position - i): count}, and s_i per even offset"""
    offs = [i for i in range(0, len(seq) - 15, 2)]
    codes, ok = [], []
    for i in offs:
        k = orc.lib().orc_make_kmer(seq, i)
        ok.append(k >= 0)
        codes.append(k if k >= 0 else 0)
    res = o.lookup(np.array(codes, dtype=np.uint32))
    votes, s = {}, []
    for i, good, (kind, n, sites) in zip(offs, ok, res):
        if not good or kind in (0, 3):      # invalid / absent / HIGH never vote
            s.append(0)
            continue
        s.append(n)
        for (c, p) in sites:
            d = (c, p - i)
            if orc.lib().orc_gp_to_i64(c, p - i) == 0:
                continue                    # packed key 0 is the "no hit" bucket (indexer.rs:337)
            votes[d] = votes.get(d, 0) + 1
    return votes, s


def test_gate_pass_implies_no_drop():
    genes = _panel()
    o = orc.OracleIndex(genes)
    rng = random.Random(9)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    n_gate = n_total = 0
    for it in range(3000):
        ga, gb = rng.randrange(len(genes)), rng.randrange(len(genes))
        L = rng.choice((100, 150, 151, 230, 250))
        cut = rng.randint(20, L - 20) if it % 3 else L
        sa = rng.randrange(0, 3000 - L)
        sb = rng.randrange(0, 3000 - L)
        seq = bytearray(genes[ga][0][sa:sa + cut] + genes[gb][0][sb:sb + L - cut])
        if rng.random() < 0.5:
            seq = bytearray(bytes(seq[::-1]).translate(comp))
        for _ in range(rng.choice((0, 0, 1, 2, 4))):
            seq[rng.randrange(len(seq))] = rng.choice(b"ACGTN")
        seq = bytes(seq)
        votes, s = _votes(o, seq)
        counts = sorted(votes.values(), reverse=True) + [0, 0]
        count1, count2 = counts[0], counts[1]
        P = sum(1 for x in s if x >= 1)
        T = sum(min(x, 2) for x in s)
        # the inequalities the screen relies on hold for every read
        assert count1 <= P and count1 + count2 <= T
        n_total += 1
        if count1 * 2 >= 40 and count2 * 2 >= 20:            # the gate of indexer.rs:353-360
            n_gate += 1
            assert P >= NEED_MAJOR and 2 * P >= NEED_MAJOR + NEED_MINOR          # k_scan keeps it
            assert T >= NEED_MAJOR + NEED_MINOR                                  # k_diag keeps it ...
            for d, c_d in votes.items():                                          # ... whatever diagonal seeded it
                assert T - c_d >= NEED_MINOR, (d, c_d, T)
    assert n_gate > 300 and n_total == 3000
    o.close()
