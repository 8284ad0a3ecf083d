"""SURVEY 8(a) row 17 + 8(f) #3: FusionMapper::add_match buckets (src/core/fusion_mapper.rs:253-275), the per-record
filters of filter_matches (:298-377), sort_matches (:379-385) with ReadMatch::partial_cmp (src/core/read_match.rs:203-229).

CPU: the oracle twin (orc_bucket_sort, a literal restatement with std::stable_sort) against the independently written Python
mirror in genefuserust_b200/host.py, fuzzed with name ties; -m gpu: records filtered and ordered on the device
(GF_OUT_DROP_FILTERED | GF_OUT_BUCKET_ORDER) + the host's name tie-break against the oracle's order of the oracle's records.
"""
import random

import pytest

import _oracle as orc
from genefuserust_b200 import host
from genefuserust_b200._abi import GF_OUT_BUCKET_ORDER, GF_OUT_DROP_FILTERED, gf_match, gf_match_order_key


def mk(pair, rb, slen, lc, rc, flags=0, source=1):
    m = gf_match()
    m.pair_idx, m.read_break, m.seq_len, m.l_contig, m.r_contig, m.filter_flags, m.source = pair, rb, slen, lc, rc, flags, source
    return m


def python_mirror(records, n_genes, names, drop):
    """buckets in index order, each sorted by host.sort_read_matches (written independently of the oracle)"""
    buckets = {}
    for i, m in enumerate(records):
        if drop and m.filter_flags:
            continue
        buckets.setdefault(host.fusion_bucket(n_genes, m), []).append(i)
    out = []
    for b in sorted(buckets):
        idx = list(buckets[b])
        wrapped = [records[i] for i in idx]
        pos = {id(r): i for r, i in zip(wrapped, idx)}
        host.sort_read_matches(wrapped, lambda m: names[pos[id(m)]])
        out.extend((pos[id(r)], b) for r in wrapped)
    return out


def test_bucket_sort_hand_case():
    names = [b"r0", b"r1", b"r2", b"r3", b"r4", b"r5", b"r5"]
    ms = [mk(0, 50, 150, 1, 2), mk(1, 70, 150, 1, 2), mk(2, 70, 120, 1, 2), mk(3, 70, 120, 1, 2), mk(4, 50, 150, 1, 2),
          mk(5, 60, 99, 1, 2), mk(6, 60, 99, 1, 2)]
    got = orc.bucket_sort(ms, 136, names)
    # break 70 first (len 120 before 150; of the two 120s the larger name first), then 60 (equal names keep push order),
    # then 50 (r4 before r0); bucket = 136 * right + left
    assert [i for i, _ in got] == [3, 2, 1, 5, 6, 4, 0]
    assert {b for _, b in got} == {136 * 2 + 1}


def test_bucket_sort_fuzz_name_ties():
    rng = random.Random(17)
    for case in range(300):
        n_genes = rng.choice([2, 5, 40])
        n = rng.randrange(0, 60)
        recs, names = [], []
        for i in range(n):
            recs.append(mk(i // 2, rng.choice([40, 41, 90]), rng.choice([100, 150, 151]), rng.randrange(n_genes),
                           rng.randrange(n_genes), flags=rng.choice([0, 0, 0, 1, 2, 4, 6]), source=i % 2 + 1))
            # few distinct names, prefixes of each other, bytes >= 0x80: String order is byte-wise
            names.append(rng.choice([b"@a", b"@a 1", b"@ab", b"@b", b"@a\xc3\xa9", b"@a merged_diff_0"]))
        for drop in (True, False):
            assert orc.bucket_sort(recs, n_genes, names, drop) == python_mirror(recs, n_genes, names, drop), (case, drop)


def test_order_key_is_the_sort_key():
    """gf_match_order_key (header, mirrored in _abi.py) orders exactly like (bucket, read_break desc, seq_len asc)"""
    rng = random.Random(3)
    recs = [mk(i, rng.randrange(0, 600), rng.randrange(30, 500), rng.randrange(136), rng.randrange(136)) for i in range(2000)]
    by_key = sorted(recs, key=lambda m: gf_match_order_key(136, m))
    by_def = sorted(recs, key=lambda m: (136 * m.r_contig + m.l_contig, -m.read_break, m.seq_len))
    assert [(m.r_contig, m.l_contig, m.read_break, m.seq_len) for m in by_key] == \
           [(m.r_contig, m.l_contig, m.read_break, m.seq_len) for m in by_def]


def read_name(batch_names, m):
    """m_read.m_name of the ReadMatch a record stands for: the merged read is named '{R1 name} merged_diff_{N}'
    (src/core/read.rs:372), an R1 / R2 match (forward or reverse complement) keeps the read's name (:243-261)"""
    n1, n2 = batch_names
    if m.source == 0:
        return n1[m.pair_idx] + b" merged_diff_%d" % m.merge_diff
    return n1[m.pair_idx] if m.source == 1 else n2[m.pair_idx]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [GF_OUT_DROP_FILTERED, GF_OUT_BUCKET_ORDER, GF_OUT_DROP_FILTERED | GF_OUT_BUCKET_ORDER])
def test_device_filter_and_bucket_order(mode):
    import __graft_entry__ as ge
    ge.build()
    from genefuserust_b200 import synth
    panel = synth.make_panel(scale=0.02)
    b = synth.generate_pairs(panel, 120000, read_len=150, seed=21, p_fusion=0.05, threads=8)
    # names with many ties: pairs share names in groups, R1 / R2 of a pair share one name in every third group
    n1 = [b"@SYN:%d 1" % (i // 7) for i in range(b.n)]
    n2 = [(b"@SYN:%d 1" if (i // 7) % 3 == 0 else b"@SYN:%d 2") % (i // 7) for i in range(b.n)]
    m = host.FusionMapper.from_gene_spans(panel.genes(), device=0)
    o = orc.OracleIndex(panel.genes())
    want_all = o.scan(b, threads=8)
    recs = []
    for t in want_all:
        r = gf_match()
        for f, v in zip(gf_match.FIELDS, t):
            setattr(r, f, v)
        recs.append(r)
    drop = bool(mode & GF_OUT_DROP_FILTERED)
    if mode & GF_OUT_BUCKET_ORDER:
        order = orc.bucket_sort(recs, m.n_genes, [read_name((n1, n2), r) for r in recs], drop)
        want = [want_all[i] for i, _ in order]
    else:
        want = [t for t in want_all if not (drop and t[15])]
    assert len(want) > 500 and (not drop or len(want) < len(want_all))
    m.set_output_mode(mode)
    got = m.scan_pair_end(b)
    if mode & GF_OUT_BUCKET_ORDER:
        got = m.finish_order(got, lambda r: read_name((n1, n2), r))
    assert [r.astuple() for r in got] == want
    st = m.map_stats()
    assert st.n_matches == len(want_all)          # the counter still says how many matches the scan made
    # chunked host path (several pipeline chunks) and the list call give the same order
    import os
    os.environ["GF_CHUNK_MB"] = "4"
    try:
        got2 = m.scan_pair_end(b)
        got3 = host.scan_list([m], b)[0]
    finally:
        del os.environ["GF_CHUNK_MB"]
    if mode & GF_OUT_BUCKET_ORDER:
        got2 = m.finish_order(got2, lambda r: read_name((n1, n2), r))
        got3 = m.finish_order(got3, lambda r: read_name((n1, n2), r))
    assert [r.astuple() for r in got2] == want and [r.astuple() for r in got3] == want
    m.set_output_mode(0)
    assert [r.astuple() for r in m.scan_pair_end(b)] == want_all
    m.close()
    o.close()
