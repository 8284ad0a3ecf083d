"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact (integer / byte / index work): every field of every match record must be identical."""
import os
import random

import numpy as np
import pytest

import _oracle as orc
from genefuserust_b200 import ReadBatch, synth
from genefuserust_b200._abi import gf_params

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def host():
    import __graft_entry__ as ge
    ge.build()
    from genefuserust_b200 import host as h
    return h


@pytest.fixture(scope="module")
def small_panel():
    return synth.make_panel(scale=0.02)


@pytest.fixture(scope="module")
def mappers(host, small_panel):
    m = host.FusionMapper.from_gene_spans(small_panel.genes(), device=0)
    o = orc.OracleIndex(small_panel.genes())
    yield m, o
    m.close()
    o.close()


def assert_same_matches(got, want, ctx=""):
    got = [m.astuple() for m in got]
    if got != want:
        sg, sw = set(got), set(want)
        only_g = sorted(sg - sw)[:5]
        only_w = sorted(sw - sg)[:5]
        raise AssertionError(f"{ctx}: {len(got)} GPU vs {len(want)} oracle records; only GPU {only_g}; only oracle {only_w}")


def test_index_counts_and_lookup(mappers):
    m, o = mappers
    info = m.m_indexer.info()
    c = o.counts()
    assert (info.n_sites, info.n_keys, info.n_unique, info.n_normal, info.n_high) == (
        c["n_sites"], c["n_keys"], c["n_unique"], c["n_normal"], c["n_high"])
    keys = o.keys()
    rng = np.random.RandomState(5)
    sample = np.concatenate([keys[rng.randint(0, len(keys), 60000)], rng.randint(0, 2**32, 20000).astype(np.uint32),
                             np.array([0, 0xFFFFFFFF, 1, 0x55555555], dtype=np.uint32)])
    got = m.m_indexer.lookup(sample)
    want = o.lookup(sample)
    bad = [(int(k), g, w) for k, g, w in zip(sample, got, want) if g != w]
    assert not bad, bad[:5]
    # every kind must have been exercised
    kinds = {w[0] for w in want}
    assert kinds == {0, 1, 2, 3}


def test_index_dedup_rule_thresholds(host):
    # closed form of index_contig's state machine: n occurrences -> unique / NORMAL (2..max(T,2)) / HIGH
    blk = synth.random_bases(99, 64).tobytes()
    genes = []
    for copies in (1, 2, 3, 5, 6, 9):
        g = b"".join(synth.random_bases(1000 + copies * 10 + k, 40).tobytes() + blk[copies:copies + 40] for k in range(copies))
        genes.append((g, False))
    for thr in (0, 1, 2, 5, 7):
        p = gf_params.default()
        p.skip_key_dup_threshold = thr
        m = host.FusionMapper.from_gene_spans(genes, params=p, device=0)
        o = orc.OracleIndex(genes, params=p)
        info = m.m_indexer.info()
        c = o.counts()
        assert (info.n_sites, info.n_keys, info.n_unique, info.n_normal, info.n_high) == (
            c["n_sites"], c["n_keys"], c["n_unique"], c["n_normal"], c["n_high"]), thr
        keys = o.keys()
        assert m.m_indexer.lookup(keys) == o.lookup(keys)
        m.close()
        o.close()


def _edge_pairs(rng):
    """hand-made pairs around fast_merge's rules: N, lower case, ragged lengths, low/high qualities"""
    def rnd(n):
        return bytes(rng.choice(b"ACGT") for _ in range(n))
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    pairs = []
    for _ in range(400):
        flen = rng.randint(20, 320)
        l1 = rng.randint(16, 160)
        l2 = rng.randint(16, 160)
        frag = bytearray(rnd(max(flen, l1, l2)))
        r1 = bytearray(frag[:l1])
        r2 = bytearray(bytes(frag[::-1]).translate(comp)[:l2])
        q1 = bytearray(rng.choice(b"EA/?0#") for _ in range(l1))
        q2 = bytearray(rng.choice(b"EA/?0#") for _ in range(l2))
        for _ in range(rng.randint(0, 4)):
            which = rng.choice((r1, r2))
            k = rng.randrange(len(which))
            which[k] = rng.choice(b"ACGTNacgtnRY.")
        pairs.append(((bytes(r1), bytes(q1)), (bytes(r2), bytes(q2))))
    return pairs


def test_fast_merge_parity(mappers, small_panel):
    m, _ = mappers
    rng = random.Random(3)
    pairs = _edge_pairs(rng)
    b = ReadBatch.from_reads([p[0] for p in pairs], [p[1] for p in pairs])
    got = m.fast_merge(b)
    n_merged = 0
    for i, (r1, r2) in enumerate(pairs):
        w = orc.fast_merge(r1[0], r1[1], r2[0], r2[1])
        want = (1, w[2], w[3], len(w[0])) if w else (0, 0, 0, 0)
        assert got[i] == want, (i, got[i], want, r1, r2)
        n_merged += bool(w)
    assert 20 < n_merged < len(pairs)
    # synthetic reads too
    b = synth.generate_pairs(small_panel, 20000, read_len=150, seed=5)
    got = m.fast_merge(b)
    for i in range(0, b.n, 7):
        s1, q1 = b.read(i, 1)
        s2, q2 = b.read(i, 2)
        w = orc.fast_merge(s1, q1, s2, q2)
        want = (1, w[2], w[3], len(w[0])) if w else (0, 0, 0, 0)
        assert got[i] == want, (i, got[i], want)


@pytest.mark.parametrize("read_len,seed", [(75, 11), (150, 12), (250, 13)])
def test_scan_pair_end_parity(mappers, small_panel, read_len, seed):
    m, o = mappers
    b = synth.generate_pairs(small_panel, 60000, read_len=read_len, seed=seed, p_fusion=0.03)
    got = m.scan_pair_end(b)
    want = o.scan(b, threads=8)
    assert len(want) > 50
    assert_same_matches(got, want, f"L={read_len}")
    st = m.map_stats()
    oc = o.counters()
    assert st.n_pairs == b.n
    assert st.n_matches == len(want)
    assert st.n_survivors >= oc["n_gated"] - 2 * len(want)  # the screen is conservative: it keeps every gated read
    assert st.n_probes_pass1 > 0 and st.kernel_launches >= 3


def test_scan_pair_end_noisy_fusions(mappers, small_panel):
    """many fusion reads with a high error rate: exercises the exact path, the mismatch gate, rc retries and
    non-zero edit distances"""
    m, o = mappers
    b = synth.generate_pairs(small_panel, 30000, read_len=150, seed=77, p_target=0.3, p_fusion=0.6, sub_rate=0.01,
                             n_rate=0.002)
    got = m.scan_pair_end(b)
    want = o.scan(b, threads=8)
    assert len(want) > 2000
    assert any(w[10] > 0 or w[11] > 0 for w in want)       # non-zero distances
    assert any(w[2] == 1 for w in want)                     # rc retries
    assert_same_matches(got, want, "noisy")


def _repeat_rich_panel():
    """genes that share long blocks (2, 3, 5 = NORMAL dupes; 6 copies = HIGH; one copy reverse-complemented), with
    fusion breakpoints inside and at the edges of the shared blocks: votes spread over several diagonals per k-mer,
    which is what the screen's bound  count1 + count2 <= sum min(sites, 2),  count1 <= #present  has to survive"""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    genes = [synth.random_bases(4000 + g, 6000) for g in range(24)]
    def put(g, pos, blk):
        genes[g][pos:pos + len(blk)] = blk
    a, b, c, d = (synth.random_bases(900 + k, n) for k, n in enumerate((1500, 1200, 800, 700)))
    for g in (0, 1, 2):
        put(g, 1000 + 500 * g, a)
    put(3, 2000, b); put(4, 700, b)
    put(5, 3000, np.frombuffer(b.tobytes()[::-1].translate(comp), dtype=np.uint8))
    for g in (6, 7, 8, 9, 10):
        put(g, 400 * (g - 5), c)
    for g in (11, 12, 13, 14, 15, 16):
        put(g, 2500, d)
    fusions = [(0, 1700, 1, 3, 2500, 1), (1, 1500, 1, 20, 3000, -1), (2, 2000, -1, 4, 1200, 1), (5, 3600, 1, 21, 1000, 1),
               (6, 600, 1, 7, 1100, -1), (8, 1500, -1, 22, 4000, -1), (11, 2800, 1, 23, 500, 1), (12, 2500, 1, 9, 1900, 1),
               (3, 1999, 1, 0, 999, 1), (10, 2400, -1, 1, 3001, 1), (17, 3000, 1, 18, 3000, 1), (19, 100, 1, 17, 5900, -1)]
    revs = [g % 3 == 0 for g in range(24)]
    return synth.Panel([f"G{g}" for g in range(24)], genes, [int(r) for r in revs], fusions)


@pytest.mark.parametrize("read_len,seed", [(150, 5), (100, 6), (250, 7)])
def test_repeat_rich_panel(host, read_len, seed):
    panel = _repeat_rich_panel()
    m = host.FusionMapper.from_gene_spans(panel.genes(), device=0)
    o = orc.OracleIndex(panel.genes())
    c = o.counts()
    assert c["n_normal"] > 3000 and c["n_high"] > 500
    b = synth.generate_pairs(panel, 40000, read_len=read_len, seed=seed, p_target=0.45, p_fusion=0.5, sub_rate=0.004,
                             n_rate=0.001)
    got = m.scan_pair_end(b)
    want = o.scan(b, threads=8)
    assert len(want) > 3000
    assert_same_matches(got, want, f"repeat-rich L={read_len}")
    se = ReadBatch(b.seq2, b.qual2, b.off2)
    assert_same_matches(m.scan_single_end(se), o.scan(se, threads=8), f"repeat-rich SE L={read_len}")
    m.close()
    o.close()


def test_scan_single_end_parity(mappers, small_panel):
    m, o = mappers
    b = synth.generate_pairs(small_panel, 40000, read_len=150, seed=21, p_fusion=0.05)
    se = ReadBatch(b.seq1, b.qual1, b.off1)
    got = m.scan_single_end(se)
    want = o.scan(se, threads=8)
    assert len(want) > 50
    assert_same_matches(got, want, "SE")


def test_edge_cases(mappers, small_panel, host):
    m, o = mappers
    # empty batch
    empty = ReadBatch(np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint64),
                      np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert m.scan_pair_end(empty) == []
    # ragged reads, lower case, N, tiny reads, empty reads, long (1000) reads built from gene sequence
    rng = random.Random(11)
    genes = small_panel.seqs
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    r1s, r2s = [], []
    for k in range(600):
        ga, gb = rng.randrange(len(genes)), rng.randrange(len(genes))
        la = rng.choice((0, 1, 15, 16, 17, 40, 100, 151, 300, 640, 1000))
        sa = rng.randrange(0, max(1, len(genes[ga]) - 600))
        sb = rng.randrange(0, max(1, len(genes[gb]) - 600))
        x = rng.randint(30, 500)
        frag = bytearray(genes[ga][sa:sa + x].tobytes() + genes[gb][sb:sb + 1100 - x].tobytes())
        if rng.random() < 0.5:
            frag = bytearray(bytes(frag[::-1]).translate(comp))
        for _ in range(rng.randint(0, 3)):
            p = rng.randrange(len(frag))
            frag[p] = rng.choice(b"acgtNnRACGT")
        l1 = la
        l2 = rng.choice((0, 16, 75, 150, 250, 1000)) if k % 3 else la
        r1 = bytes(frag[:l1])
        r2 = bytes(frag[:max(l1, l2, 1) + rng.randint(0, 60)][::-1]).translate(comp)[:l2]
        q = lambda n: bytes(rng.choice(b"EEEEEEA/") for _ in range(n))
        r1s.append((r1, q(len(r1))))
        r2s.append((r2, q(len(r2))))
    b = ReadBatch.from_reads(r1s, r2s)
    assert b.max_len == 1000
    got = m.scan_pair_end(b)
    want = o.scan(b, threads=4)
    assert len(want) > 10
    if o.counters()["n_panic"] == 0:
        assert m.last_rc == 0
    else:
        assert m.last_rc == -5  # GF_E_REF_PANIC: > 640 columns on both sides, records still exact
    assert_same_matches(got, want, "edge")


def test_max_len_hint_is_enforced(mappers, small_panel, host):
    m, _ = mappers
    b = synth.generate_pairs(small_panel, 1000, read_len=300, seed=1)
    b.max_len = 150  # lie: reads are longer than the hint -> must fail loudly, never truncate
    with pytest.raises(host.GeneFuseError):
        m.scan_pair_end(b)


def test_missing_max_len_hint(mappers, small_panel):
    """max_len = 0 (no hint): gf_map_pairs finds the longest read itself; same records, and the thread-per-pair kernels
    (zero-copy qualities are only offered by them) still run"""
    import torch
    m, o = mappers
    b = synth.generate_pairs(small_panel, 30000, read_len=150, seed=33, p_fusion=0.1)
    want = [r.astuple() for r in m.scan_pair_end(b)]
    pinned = [torch.from_numpy(a.copy()).pin_memory() for a in (b.seq1, b.qual1, b.seq2, b.qual2)]
    pb = ReadBatch(pinned[0].numpy(), pinned[1].numpy(), b.off1, pinned[2].numpy(), pinned[3].numpy(), b.off2)
    pb.max_len = 0
    assert [r.astuple() for r in m.scan_pair_end(pb)] == want
    assert m.map_stats().zero_copy_qual == 1
    assert want == o.scan(b, threads=8)


def test_capacity_protocol(mappers, small_panel):
    import ctypes as C
    from genefuserust_b200._abi import gf_match
    m, o = mappers
    b = synth.generate_pairs(small_panel, 20000, read_len=150, seed=31, p_fusion=0.2)
    want = o.scan(b, threads=8)
    st = b.as_struct()
    out = (gf_match * 4)()
    n = C.c_uint64(0)
    rc = m.lib.gf_map_pairs(m.m_indexer.h, C.byref(st), out, 4, C.byref(n))
    assert rc == -3 and n.value == len(want)


def test_testdata_config1(host):
    """BASELINE config 1: testdata R1/R2 vs tinyref.fa + fusions.csv -> no gene resolves, zero matches
    ('found 0 fusions')."""
    td = os.path.join(GOLD, "testdata")
    sc = host.PairEndScanner(os.path.join(td, "fusions.csv"), os.path.join(td, "tinyref.fa"),
                             os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq"))
    matches = sc.scan()
    assert matches == []
    info = sc.mapper.m_indexer.info()
    assert info.n_keys == 0 and len(sc.mapper.m_indexer.m_fusion_seq) == 4
    merged = sc.mapper.fast_merge(host.FastqReaderPair(os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")).read_all()[1])
    assert [x[3] for x in merged] == [178, 161, 161]
    sc.mapper.close()


def test_golden_fixture_matches(host):
    """the committed golden records (tests/golden/synth_small_matches.json) through the CUDA path"""
    import hashlib
    import json
    g = json.load(open(os.path.join(GOLD, "synth_small_matches.json")))
    panel = synth.make_panel(scale=g["panel_scale"], max_genes=g["max_genes"])
    batch = synth.generate_pairs(panel, g["n_pairs"], read_len=g["read_len"], seed=g["seed"], p_fusion=g["p_fusion"],
                                 threads=2)
    assert hashlib.sha256(batch.seq1.tobytes() + batch.seq2.tobytes()).hexdigest() == g["reads_sha256"]
    m = host.FusionMapper.from_gene_spans(panel.genes(), device=0)
    info = m.m_indexer.info()
    assert {"n_sites": info.n_sites, "n_keys": info.n_keys, "n_unique": info.n_unique, "n_normal": info.n_normal,
            "n_high": info.n_high} == g["index_counts"]
    got = [list(r.astuple()) for r in m.scan_pair_end(batch)]
    assert got == g["matches"]
    m.close()


def test_full_size_properties(host):
    """BASELINE-size check: 4 M pairs of the bench workload against the full-size panel — the whole record set is
    compared with the oracle (it is tiny), plus size-independent properties: mapping twice gives the same records,
    mapping two halves separately gives the union, every record satisfies the structural invariants."""
    panel = synth.make_panel(scale=1.0)
    n = 4_000_000
    b = synth.generate_pairs(panel, n, read_len=150, seed=12)
    m = host.FusionMapper.from_gene_spans(panel.genes(), device=0)
    got = [r.astuple() for r in m.scan_pair_end(b)]
    assert got == [r.astuple() for r in m.scan_pair_end(b)]                       # idempotent
    half = n // 2
    lo = [r.astuple() for r in m.scan_pair_end(b.slice(0, half))]
    hi = [r.astuple() for r in m.scan_pair_end(b.slice(half, n))]
    assert got == lo + [(r[0] + half,) + r[1:] for r in hi]                       # shard-invariant
    assert got == sorted(got, key=lambda r: (r[0], r[1]))                         # (pair_idx, source) order
    for r in got:
        pair, source, used_rc, reversed_, rb, lc, lp, rc, rp, gap, ld, rd, slen, olen, diff, ff = r
        assert (ff & 2) == (2 if ld + rd >= 5 else 0) and ff < 8
        assert 0 <= pair < n and source in (0, 1, 2) and 0 <= rb < slen
        assert reversed_ == (1 if (used_rc and source != 0) else 0)
        assert (olen >= 30) == (source == 0) and 0 <= diff <= 2
        assert 0 <= lc < panel.n_genes and 0 <= rc < panel.n_genes
    o = orc.OracleIndex(panel.genes())
    want = o.scan(b, threads=os.cpu_count() or 8)
    assert len(want) > 1000
    assert got == want
    m.close()
    o.close()


def test_zero_copy_qualities_pinned(mappers, small_panel):
    """pinned host arenas: qualities are not copied, the kernels fetch the few bytes fast_merge needs over PCIe;
    results must be identical to the staged path and to the oracle"""
    import torch
    m, o = mappers
    b = synth.generate_pairs(small_panel, 50000, read_len=150, seed=41, p_fusion=0.05, sub_rate=0.01)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    bp = ReadBatch(pin(b.seq1), pin(b.qual1), pin(b.off1.view(np.int64)).view(np.uint64),
                   pin(b.seq2), pin(b.qual2), pin(b.off2.view(np.int64)).view(np.uint64))
    want = o.scan(b, threads=8)
    got_staged = m.scan_pair_end(b)
    assert m.map_stats().zero_copy_qual == 0
    got_zc = m.scan_pair_end(bp)
    st = m.map_stats()
    assert st.zero_copy_qual == 1 and st.h2d_bytes < 0.6 * (4 * b.n * 150)
    assert_same_matches(got_staged, want, "staged")
    assert_same_matches(got_zc, want, "zero-copy")


def test_packed_upload_parity(mappers, small_panel, monkeypatch):
    """packed upload (csrc/gf_pack.cpp): with all four arenas in pinned memory the host threads build the reads' bit-planes and
    only those are copied; k_prep takes them as they are (mate 2 reversed on the device), survivors are read from the mapped
    arenas.  Ragged reads (0 .. 256 bases) with N / lower case / IUPAC bytes, PE and SE, several thread counts: the records must
    equal the oracle's and the ASCII path's."""
    import torch
    m, o = mappers
    rng = random.Random(5)
    genes = small_panel.seqs
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    r1s, r2s = [], []
    for k in range(40000):
        ga = rng.randrange(len(genes))
        sa = rng.randrange(0, max(1, len(genes[ga]) - 700))
        if k % 4 == 0:   # fusion-like fragment
            gb = rng.randrange(len(genes))
            sb = rng.randrange(0, max(1, len(genes[gb]) - 700))
            x = rng.randint(40, 300)
            frag = bytearray(genes[ga][sa:sa + x].tobytes() + genes[gb][sb:sb + 600 - x].tobytes())
        else:
            frag = bytearray(genes[ga][sa:sa + 600].tobytes())
        if rng.random() < 0.5:
            frag = bytearray(bytes(frag[::-1]).translate(comp))
        flen = rng.randint(60, 500)
        frag = frag[:flen]
        if k % 3 == 0:
            for _ in range(rng.randint(1, 4)):
                frag[rng.randrange(len(frag))] = rng.choice(b"acgtNnRYK@\x00\xffACGT")
        l1 = rng.choice((0, 1, 15, 16, 31, 32, 33, 64, 65, 100, 150, 151, 200, 256)) if k % 5 == 0 else 150
        l2 = rng.choice((0, 16, 75, 150, 250, 256)) if k % 7 == 0 else l1
        r1 = bytes(frag[:l1])
        r2 = bytes(frag[::-1]).translate(comp)[:l2]
        q = lambda n: bytes(rng.choice(b"EEEEEEA/") for _ in range(n))
        r1s.append((r1, q(len(r1))))
        r2s.append((r2, q(len(r2))))
    b = ReadBatch.from_reads(r1s, r2s)
    assert b.max_len == 256
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    bp = ReadBatch(pin(b.seq1), pin(b.qual1), pin(b.off1.view(np.int64)).view(np.uint64),
                   pin(b.seq2), pin(b.qual2), pin(b.off2.view(np.int64)).view(np.uint64))
    bp.max_len = b.max_len
    want = o.scan(b, threads=8)
    assert len(want) > 300
    monkeypatch.setenv("GF_HOST_PACK", "0")
    got_ascii = m.scan_pair_end(bp)
    assert m.map_stats().packed_upload == 0
    assert_same_matches(got_ascii, want, "ascii upload")
    monkeypatch.setenv("GF_HOST_PACK", "1")
    monkeypatch.setenv("GF_CHUNK_MB", "1")       # many pipeline chunks
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("GF_PACK_THREADS", threads)
        got = m.scan_pair_end(bp)
        st = m.map_stats()
        if st.packed_upload == 0:
            pytest.skip("host without AVX-512BW: the packed upload is not offered")
        assert st.h2d_bytes < 0.5 * (b.seq1.size + b.seq2.size)
        assert_same_matches(got, want, f"packed upload, {threads} threads")
    # a read longer than the hint: the packing threads notice (they read the offsets anyway) and the call fails loudly
    bp.max_len = 200
    with pytest.raises(Exception):
        m.scan_pair_end(bp)
    bp.max_len = b.max_len
    # single end
    se = ReadBatch(bp.seq1, bp.qual1, bp.off1, None, None, None)
    se.max_len = b.max_len
    want_se = o.scan(ReadBatch(b.seq1, b.qual1, b.off1, None, None, None), threads=8)
    got_se = m.scan_single_end(se)
    assert m.map_stats().packed_upload == 1
    assert_same_matches(got_se, want_se, "packed upload, single end")


def test_packed_upload_uniform_chunks(mappers, small_panel, monkeypatch):
    """packed upload, compact form: a chunk whose reads all have one length sends no offset tables (the device fills them,
    k_fill_uniform) and no exception table unless a read of the chunk has a byte that is not upper-case ACGT — and then only
    the packing threads that met one wrote their part of it.  2x100 reads, exceptions in one narrow region of the batch only,
    every chunk packed / the hybrid mix, several thread counts and chunk sizes: records equal the oracle's."""
    import torch
    m, o = mappers
    rng = random.Random(11)
    genes = small_panel.seqs
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    n = 60000
    r1s, r2s = [], []
    for k in range(n):
        ga = rng.randrange(len(genes))
        sa = rng.randrange(0, max(1, len(genes[ga]) - 400))
        if k % 4 == 0:
            gb = rng.randrange(len(genes))
            sb = rng.randrange(0, max(1, len(genes[gb]) - 400))
            x = rng.randint(40, 200)
            frag = bytearray(genes[ga][sa:sa + x].tobytes() + genes[gb][sb:sb + 300 - x].tobytes())
        else:
            frag = bytearray(genes[ga][sa:sa + 300].tobytes())
        frag = frag[:rng.randint(100, 300)]
        if len(frag) < 100:
            frag = frag + bytearray(b"A" * (100 - len(frag)))
        if 20000 <= k < 20500 and k % 2 == 0:   # the only flagged reads of the batch
            frag[rng.randrange(100)] = rng.choice(b"acgtNn@")
            frag[len(frag) - 1 - rng.randrange(100)] = rng.choice(b"acgtNn@")
        r1 = bytes(frag[:100])
        r2 = bytes(frag[::-1]).translate(comp)[:100]
        q = lambda c: bytes(rng.choice(b"EEEEEEA/") for _ in range(c))
        r1s.append((r1, q(100)))
        r2s.append((r2, q(100)))
    b = ReadBatch.from_reads(r1s, r2s)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    bp = ReadBatch(pin(b.seq1), pin(b.qual1), pin(b.off1.view(np.int64)).view(np.uint64),
                   pin(b.seq2), pin(b.qual2), pin(b.off2.view(np.int64)).view(np.uint64))
    bp.max_len = 100
    want = o.scan(b, threads=8)
    assert len(want) > 300
    for force, chunk_mb, threads, parts in (("1", "1", "5", None), ("1", "8", "16", "64"), ("1", "64", "3", "1"), (None, "1", "8", None),
                                            (None, "2", "12", "7")):
        if parts is None:      # parts per packing job (default: one per thread): the output layout depends on them, the records must not
            monkeypatch.delenv("GF_PACK_PARTS", raising=False)
        else:
            monkeypatch.setenv("GF_PACK_PARTS", parts)
        if force is None:
            monkeypatch.delenv("GF_HOST_PACK", raising=False)
            monkeypatch.setenv("GF_PACK_MIN_THREADS", "1")
        else:
            monkeypatch.setenv("GF_HOST_PACK", force)
        monkeypatch.setenv("GF_CHUNK_MB", chunk_mb)
        monkeypatch.setenv("GF_PACK_THREADS", threads)
        got = m.scan_pair_end(bp)
        st = m.map_stats()
        if st.packed_upload == 0 and force == "1":
            pytest.skip("host without AVX-512BW: the packed upload is not offered")
        if force == "1":
            assert st.packed_upload == 1
            assert st.h2d_bytes < 0.4 * (b.seq1.size + b.seq2.size)   # plane words only (32 bytes per 100-base read) + the few exception words
        assert_same_matches(got, want, f"packed upload, uniform reads, force={force} chunk={chunk_mb} MB threads={threads}")


@pytest.mark.parametrize("pinned", [False, True])
def test_list_mode_concurrent_handles(host, small_panel, pinned, monkeypatch):
    """multi-CSV list mode (fusion_scan.rs:62-188): one index per CSV, used concurrently from different host
    threads over the same reads; handles are independent, every result must equal the oracle's for its panel.
    pinned: the arenas are pinned host memory, so the calls upload chunk by chunk as planes or ASCII and share the process's
    packing threads (a handle that finds them busy sends its chunk as ASCII)"""
    import threading
    genes = small_panel.genes()
    panels = [genes, genes[:40], genes[20:90], genes[::2]]
    b = synth.generate_pairs(small_panel, 40000, read_len=150, seed=52, p_fusion=0.05)
    if pinned:
        import torch
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        b = ReadBatch(pin(b.seq1), pin(b.qual1), pin(b.off1.view(np.int64)).view(np.uint64),
                      pin(b.seq2), pin(b.qual2), pin(b.off2.view(np.int64)).view(np.uint64))
        monkeypatch.setenv("GF_CHUNK_MB", "1")
        monkeypatch.setenv("GF_PACK_THREADS", "8")
    mappers = [host.FusionMapper.from_gene_spans(p, device=0) for p in panels]
    results = [None] * len(panels)

    def work(k):
        for _ in range(3):
            results[k] = [r.astuple() for r in mappers[k].scan_pair_end(b)]
    ths = [threading.Thread(target=work, args=(k,)) for k in range(len(panels))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for k, p in enumerate(panels):
        o = orc.OracleIndex(p)
        want = o.scan(b, threads=8)
        assert results[k] == want, k
        o.close()
    assert len({len(r) for r in results}) > 1   # the panels really give different answers
    for m in mappers:
        m.close()


def _fastq_reader_reference(text):
    """FastqReader::read restated (src/core/fastq_reader.rs:75-147): four read_line calls per record, each line loses
    one trailing '\\n'; a record is returned only if all four reads returned > 0 bytes"""
    recs = []
    pos = 0

    def read_line():
        nonlocal pos
        if pos >= len(text):
            return None
        k = text.find(b"\n", pos)
        line = text[pos:] if k < 0 else text[pos:k + 1]
        pos += len(line)
        return line[:-1] if line.endswith(b"\n") else line
    while True:
        four = [read_line() for _ in range(4)]
        if any(x is None for x in four):
            break
        recs.append((four[1], four[3]))
    return recs


def _to_fastq(batch, mate, rng, eol=b"\n", final_newline=True, extra_tail=b""):
    out = []
    for i in range(batch.n):
        s, q = batch.read(i, mate)
        name = b"@SYN:%d:%d %d:N:0:ACGT" % (i, rng.randint(0, 99999), mate)
        out.append(name + eol + s + eol + b"+" + eol + q + eol)
    text = b"".join(out)
    if not final_newline and text.endswith(b"\n"):
        text = text[:-1]
    return text + extra_tail


@pytest.mark.parametrize("variant", ["plain", "no_final_newline", "crlf", "truncated_tail"])
def test_fastq_ingest_parity(mappers, small_panel, variant):
    """device FASTQ record splitting (gf_map_fastq) vs the reference reader's semantics + the oracle"""
    m, o = mappers
    rng = random.Random(9)
    b = synth.generate_pairs(small_panel, 20000, read_len=150, seed=61, p_fusion=0.05)
    kw = {}
    if variant == "no_final_newline":
        kw = dict(final_newline=False)
    elif variant == "crlf":
        kw = dict(eol=b"\r\n")          # '\r' is NOT stripped by the reference: it stays in sequence and quality
    elif variant == "truncated_tail":
        kw = dict(extra_tail=b"@incomplete\nACGTACGT\n+\n")   # dropped at EOF
    fq1, fq2 = _to_fastq(b, 1, rng, **kw), _to_fastq(b, 2, rng, **kw)
    r1, r2 = _fastq_reader_reference(fq1), _fastq_reader_reference(fq2)
    n = min(len(r1), len(r2))
    assert n == b.n
    ref_batch = ReadBatch.from_reads(r1[:n], r2[:n])
    want = o.scan(ref_batch, threads=8)
    got, nrec = m.scan_fastq(fq1, fq2)
    assert nrec == n
    assert_same_matches(got, want, variant)
    if variant != "crlf":
        assert len(want) > 50
    # single end
    got_se, nrec = m.scan_fastq(fq1, None)
    want_se = o.scan(ReadBatch.from_reads(r1), threads=8)
    assert nrec == len(r1)
    assert_same_matches(got_se, want_se, variant + " SE")


def test_fastq_ingest_testdata(host):
    """the reference's own testdata R1.fq / R2.fq through the device FASTQ path: 3 records, 0 matches, merge lengths"""
    td = os.path.join(GOLD, "testdata")
    fq1, fq2 = open(os.path.join(td, "R1.fq"), "rb").read(), open(os.path.join(td, "R2.fq"), "rb").read()
    m = host.FusionMapper.from_ref_and_fusion_files(os.path.join(td, "tinyref.fa"), os.path.join(td, "fusions.csv"))
    got, nrec = m.scan_fastq(fq1, fq2)
    assert (got, nrec) == ([], 3)
    assert len(_fastq_reader_reference(fq1)) == 3
    m.close()


def test_filter_flags(host):
    """filter_matches predicates (fusion_mapper.rs:298-377) as flags on the records: low complexity, distance >= 5,
    indel; crafted so that every flag occurs"""
    rng = random.Random(4)
    rnd = lambda n: bytes(rng.choice(b"ACGT") for _ in range(n))
    lowc = b"A" * 12 + b"C" * 13 + b"G" * 14 + b"T" * 15 + b"A" * 16          # 70 bases, 4 changes, unique 16-mers
    gene_a = rnd(400) + lowc + rnd(400)
    gene_b = rnd(1000)
    genes = [(gene_a, False), (gene_b, False)]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for k in range(400):
        kind = k % 4
        if kind == 0:      # clean fusion B | A-lowc : right side is low complexity
            frag = gene_b[300:400] + gene_a[400:400 + 50]
        elif kind == 1:    # fusion with many substitutions -> distance flag
            frag = bytearray(gene_b[100:180] + gene_a[100:170])
            for _ in range(6):
                frag[rng.randrange(20, 130)] = rng.choice(b"ACGT")
            frag = bytes(frag)
        elif kind == 2:    # same-gene "fusion" 30 bases apart -> indel flag
            frag = gene_b[500:575] + gene_b[605:680]
        else:              # plain fusion
            frag = gene_a[50:125] + gene_b[600:675]
        if rng.random() < 0.5:
            frag = frag[::-1].translate(comp)
        reads.append((frag, b"E" * len(frag)))
    b = ReadBatch.from_reads(reads)
    m = host.FusionMapper.from_gene_spans(genes, device=0)
    o = orc.OracleIndex(genes)
    got = m.scan_single_end(b)
    want = o.scan(b, threads=4)
    assert_same_matches(got, want, "flags")
    flags = {w[15] for w in want}
    assert any(f & 1 for f in flags) and any(f & 2 for f in flags) and any(f & 4 for f in flags) and 0 in flags, flags
    m.close()
    o.close()


def test_multi_device_handle(host, small_panel):
    """gf_multi_*: shards of one batch mapped by several handles from one process (device 0 listed three times on a
    1-GPU box, all visible devices otherwise); the gathered records must equal the single-handle result"""
    import torch
    genes = small_panel.genes()
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    b = synth.generate_pairs(small_panel, 100001, read_len=150, seed=71, p_fusion=0.03)
    mm = host.MultiGpuMapper(genes, devices)
    got = [r.astuple() for r in mm.scan(b)]
    o = orc.OracleIndex(genes)
    want = o.scan(b, threads=8)
    assert len(want) > 100 and got == want
    mm.close()
    o.close()


def test_adjust_fusion_break_parity(mappers, small_panel):
    """FusionResult::adjust_fusion_break on the device vs the oracle — real matches of a noisy fusion
    batch grouped per gene pair like cluster_matches does, references built with get_ref_seq like make_reference
    (fusion_result.rs:242-297), plus breaks pushed to the edges of the reads and empty references."""
    m, o = mappers
    b = synth.generate_pairs(small_panel, 30000, read_len=150, seed=78, p_target=0.3, p_fusion=0.6, sub_rate=0.01, n_rate=0.002)
    recs = o.scan(b, threads=8)
    assert len(recs) > 2000
    genes = [g for g, _ in small_panel.genes()]
    groups = {}
    for r in recs:
        pair, source, used_rc, _rev, rb, lc, lp, rc_, rp, *_ = r
        s1, q1 = b.read(pair, 1)
        s2, q2 = b.read(pair, 2)
        seq = orc.fast_merge(s1, q1, s2, q2)[0] if source == 0 else (s1 if source == 1 else s2)
        if used_rc:
            seq = orc.reverse_complement(seq)
        groups.setdefault((lc, rc_), []).append((seq, rb, lp, rp))
    results, rng = [], random.Random(4)
    for (lc, rc_), ms in sorted(groups.items()):
        # calc_fusion_point's "first match" flavour + make_reference
        lp, rp = ms[0][2], ms[0][3]
        longest_left = max(rb + 1 for _s, rb, _l, _r in ms)
        longest_right = max(len(sq) - (rb + 1) for sq, rb, _l, _r in ms)
        lref = orc.get_ref_seq(genes[lc], lp - longest_left + 1, lp)
        rref = orc.get_ref_seq(genes[rc_], rp, rp + longest_right - 1)
        jobs = [(sq, rb) for sq, rb, _l, _r in ms]
        jobs += [(sq, rng.choice((0, 2, 3, len(sq) - 4, len(sq) - 2, len(sq) + 5))) for sq, _rb, _l, _r in ms[:2]]   # edges
        results.append((lref, rref, jobs))
    results.append((b"", b"", [(results[0][2][0][0], 60)]))
    got = m.adjust_fusion_break(results)
    n_jobs = n_shift = n_undef = 0
    for (lref, rref, jobs), g in zip(results, got):
        for (sq, rb), gg in zip(jobs, g):
            want = orc.adjust_fusion_break(sq, rb, lref, rref)
            assert gg == want, (rb, len(sq), len(lref), len(rref), gg, want)
            n_jobs += 1
            n_shift += want[0] != 0
            n_undef += want[3]
    assert n_jobs > 2000 and n_shift > 20 and n_undef > 5
    assert m.last_rc == -5      # GF_E_REF_PANIC: some breaks were pushed outside their reads on purpose


def test_list_mode_shared_prep(host, small_panel):
    """BASELINE config 4 / fusion_scan.rs:62-188: one batch against several indices in one call (upload and k_prep once):
    every per-index record list must equal the one that index produces on its own, and the oracle's."""
    genes = small_panel.genes()
    subsets = [genes, genes[:40], genes[20:90], genes[:3]]
    mappers = [host.FusionMapper.from_gene_spans(g, device=0) for g in subsets]
    for L, seed in ((150, 41), (250, 42), (300, 43)):          # split screen (W = 5, 8) and the long-read path
        b = synth.generate_pairs(small_panel, 60000, read_len=L, seed=seed, p_fusion=0.1)
        got = host.scan_list(mappers, b)
        for g, m, k in zip(subsets, mappers, range(len(subsets))):
            alone = [r.astuple() for r in m.scan_pair_end(b)]
            assert [r.astuple() for r in got[k]] == alone, (L, k)
        o = orc.OracleIndex(subsets[1])
        assert [r.astuple() for r in got[1]] == o.scan(b, threads=8)
        o.close()
        assert len(got[0]) > 100
    se = ReadBatch(b.seq1, b.qual1, b.off1)
    got = host.scan_list(mappers[:2], se)
    assert [r.astuple() for r in got[1]] == [r.astuple() for r in mappers[1].scan_single_end(se)]
    for m in mappers:
        m.close()


def _device_batch(torch, b):
    """gf_batch whose pointers are DEVICE pointers (torch owns the memory); returns (struct, keepalive)"""
    from genefuserust_b200._abi import gf_batch
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (b.seq1, b.qual1, b.seq2, b.qual2)]
    o1 = torch.from_numpy(b.off1.view(np.int64)).cuda()
    o2 = torch.from_numpy(b.off2.view(np.int64)).cuda()
    db = gf_batch()
    db.n = b.n
    db.seq1, db.qual1, db.seq2, db.qual2 = (t.data_ptr() for t in d)
    db.off1, db.off2 = o1.data_ptr(), o2.data_ptr()
    db.bytes1, db.bytes2 = int(b.off1[-1]), int(b.off2[-1])
    db.max_len = b.max_len
    return db, (d, o1, o2)


def _device_records(torch, d_out, d_n, k=0):
    import ctypes as C
    from genefuserust_b200._abi import gf_match
    n = int(d_n[k].item())
    raw = d_out[:n * C.sizeof(gf_match)].cpu().numpy().tobytes()
    recs = (gf_match * n).from_buffer_copy(raw)
    return sorted((r.astuple() for r in recs), key=lambda t: (t[0], t[1]))


@pytest.mark.parametrize("read_len,seed", [(150, 51), (250, 52), (300, 53)])
def test_device_batch_entry_points(host, small_panel, read_len, seed):
    """gf_map_pairs_device / gf_map_pairs_device_list on device-resident arenas: one chunk and many chunks
    (GF_DEVICE_CHUNK_PAIRS) give the host path's records; a host-path call issued right behind an unsynchronised device call
    on the same handle waits for it (the handle's workspace is shared)"""
    import ctypes as C
    import torch
    from genefuserust_b200._abi import gf_map_stats, gf_match
    genes = small_panel.genes()
    mappers = [host.FusionMapper.from_gene_spans(g, device=0) for g in (genes, genes[:40], genes[10:70])]
    b = synth.generate_pairs(small_panel, 50000, read_len=read_len, seed=seed, p_fusion=0.1)
    want = [[r.astuple() for r in m.scan_pair_end(b)] for m in mappers]
    assert len(want[0]) > 100
    lib = mappers[0].lib
    db, keep = _device_batch(torch, b)
    cap = 2 * b.n
    K = len(mappers)
    d_outs = [torch.empty(cap * C.sizeof(gf_match), dtype=torch.uint8, device="cuda") for _ in range(K)]
    d_ns = torch.zeros(K, dtype=torch.int64, device="cuda")
    side = torch.cuda.Stream()
    for chunk_env in (None, "7000"):
        if chunk_env:
            os.environ["GF_DEVICE_CHUNK_PAIRS"] = chunk_env
        try:
            # single-index call on a side stream, then (no synchronisation) a host-path call on the same handle
            rc = lib.gf_map_pairs_device(mappers[0].m_indexer.h, C.byref(db), d_outs[0].data_ptr(), cap, d_ns.data_ptr(),
                                         C.c_void_p(side.cuda_stream))
            assert rc == 0, lib.gf_last_error()
            again = [r.astuple() for r in mappers[0].scan_pair_end(b)]
            side.synchronize()
            assert again == want[0]
            assert _device_records(torch, d_outs[0], d_ns, 0) == want[0]
            st = gf_map_stats()
            # two device calls back to back on different streams of one handle
            rc = lib.gf_map_pairs_device(mappers[0].m_indexer.h, C.byref(db), d_outs[0].data_ptr(), cap, d_ns.data_ptr(),
                                         C.c_void_p(side.cuda_stream))
            rc2 = lib.gf_map_pairs_device(mappers[0].m_indexer.h, C.byref(db), d_outs[1].data_ptr(), cap, d_ns.data_ptr() + 8,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0 and rc2 == 0, lib.gf_last_error()
            torch.cuda.synchronize()
            assert _device_records(torch, d_outs[0], d_ns, 0) == want[0] and _device_records(torch, d_outs[1], d_ns, 1) == want[0]
            assert lib.gf_get_map_stats(mappers[0].m_indexer.h, C.byref(st)) == 0
            assert st.n_matches == len(want[0]) and st.n_pairs == b.n and st.n_survivors >= len(want[0])
            # list call
            hs = (C.c_void_p * K)(*[m.m_indexer.h.value for m in mappers])
            outs = (C.c_void_p * K)(*[t.data_ptr() for t in d_outs])
            nouts = (C.c_void_p * K)(*[d_ns.data_ptr() + 8 * k for k in range(K)])
            rc = lib.gf_map_pairs_device_list(hs, K, C.byref(db), outs, cap, nouts, C.c_void_p(side.cuda_stream))
            assert rc == 0, lib.gf_last_error()
            side.synchronize()
            for k in range(K):
                assert _device_records(torch, d_outs[k], d_ns, k) == want[k], (chunk_env, k)
        finally:
            os.environ.pop("GF_DEVICE_CHUNK_PAIRS", None)
    for m in mappers:
        m.close()


@pytest.mark.parametrize("gz", [False, True, "bgzf"])
def test_fastq_stream_chunked_and_gzip(mappers, small_panel, host, gz, tmp_path):
    """the two FASTQ files as byte streams (gf_fastq_stream_*): pieces that end anywhere (mid-line, mid-record,
    different piece sizes per mate), small device chunks with the tail carried over, plain, multi-member gzip and BGZF (whose
    whole members cross PCIe compressed and are inflated on the device, one warp per member — GF_BGZF_DEVICE=0: side by side by
    all host threads —, the cut ones by the streaming decoder); the records
    must equal one gf_map_fastq call on the whole text and the oracle on the reference reader's records"""
    import gzip
    m, o = mappers
    rng = random.Random(19)
    b = synth.generate_pairs(small_panel, 30000, read_len=150, seed=71, p_fusion=0.05)
    fq1 = _to_fastq(b, 1, rng, final_newline=False)
    fq2 = _to_fastq(b, 2, rng, extra_tail=b"@incomplete\nACGT\n+\n")
    r1, r2 = _fastq_reader_reference(fq1), _fastq_reader_reference(fq2)
    n = min(len(r1), len(r2))
    want = o.scan(ReadBatch.from_reads(r1[:n], r2[:n]), threads=8)
    whole, nrec = m.scan_fastq(fq1, fq2)
    assert nrec == n and [r.astuple() for r in whole] == want and len(want) > 50

    def enc(data, members):
        if not gz:
            return data
        if gz == "bgzf":       # blocked gzip: the members that lie whole in a fed piece are inflated by all host threads
            return host.bgzf_compress(data, block=rng.choice((4000, 65280)), level=rng.choice((1, 6, 9)))
        cuts = sorted(rng.sample(range(1, len(data)), members - 1)) if members > 1 else []
        parts = [data[a:c] for a, c in zip([0] + cuts, cuts + [len(data)])]
        return b"".join(gzip.compress(p, compresslevel=1) for p in parts)     # several members, cut anywhere
    e1, e2 = enc(fq1, 3), enc(fq2, 1)
    for chunk_bytes, piece1, piece2 in ((1 << 20, 300_001, 77_777), (4096, 1 << 16, 1 << 20), (0, len(e1), len(e2))):
        st = host.FastqStream(m, paired=True, gz=gz, chunk_bytes=chunk_bytes)
        p1 = p2 = 0
        got = []
        while p1 < len(e1) or p2 < len(e2):
            st.feed(e1[p1:p1 + piece1], e2[p2:p2 + piece2])
            p1 += piece1
            p2 += piece2
            if rng.random() < 0.3:
                got += st.take()          # records may be collected at any time
        st.finish()
        got += st.take()
        recs, text_bytes, calls = st.counts()
        st.close()
        assert recs == n
        assert sorted(r.astuple() for r in got) == want, (gz, chunk_bytes)
        if chunk_bytes and chunk_bytes <= (1 << 20):
            assert calls > 3
    # files on disk, format by extension (FastqReader::new), single end too
    ext = ".fq.gz" if gz else ".fq"
    f1, f2 = str(tmp_path / ("a_R1" + ext)), str(tmp_path / ("a_R2" + ext))
    open(f1, "wb").write(e1)
    open(f2, "wb").write(e2)
    got, (recs, _tb, _calls) = host.FastqStream.scan_files(m, f1, f2, piece=1 << 18, chunk_bytes=1 << 21)
    assert recs == n and [r.astuple() for r in got] == want
    got, (recs, _tb, _calls) = host.FastqStream.scan_files(m, f1, None, piece=1 << 18, chunk_bytes=1 << 21)
    assert recs == len(r1) and [r.astuple() for r in got] == o.scan(ReadBatch.from_reads(r1), threads=8)
    if gz:      # a truncated gzip file is an error, not a silent short read
        st = host.FastqStream(m, paired=False, gz=True)
        st.feed(e1[:len(e1) // 2])
        with pytest.raises(host.GeneFuseError):
            st.finish()
        st.close()
    if gz == "bgzf":   # a damaged member (inflated on the device, one warp per member) is an error as well: sizes and CRC-32 are checked
        bad = bytearray(e1)
        for pos in (len(bad) // 3, len(bad) // 2 + 17):
            bad[pos] ^= 0x10
        st = host.FastqStream(m, paired=False, gz=True)
        with pytest.raises(host.GeneFuseError):
            st.feed(bytes(bad))
            st.finish()
        st.close()


def test_pack_stream_batched_shim(mappers, small_panel, host):
    """gf_stream_*: packs of 1000 pairs (the reference's granularity, common.rs:23) pushed out of order with the caller's pair
    numbering, mapped in batches; same records as one gf_map_pairs call over all pairs"""
    m, o = mappers
    b = synth.generate_pairs(small_panel, 24500, read_len=150, seed=81, p_fusion=0.05)
    want = o.scan(b, threads=8)
    assert len(want) > 50
    packs = [(lo, min(lo + 1000, b.n)) for lo in range(0, b.n, 1000)]
    random.Random(5).shuffle(packs)
    for batch_pairs in (4000, 0):
        st = host.PackStream(m, paired=True, batch_pairs=batch_pairs)
        for lo, hi in packs:
            st.push(lo, [b.read(i, 1) for i in range(lo, hi)], [b.read(i, 2) for i in range(lo, hi)])
        st.flush()
        got = st.take()
        pushed, calls = st.counts()
        st.close()
        assert pushed == b.n and calls == (7 if batch_pairs else 1)
        assert [r.astuple() for r in got] == want
    st = host.PackStream(m, paired=False, batch_pairs=3000)
    for lo, hi in packs:
        st.push(lo, [b.read(i, 1) for i in range(lo, hi)])
    st.flush()
    se = ReadBatch(b.seq1, b.qual1, b.off1)
    assert [r.astuple() for r in st.take()] == o.scan(se, threads=8)
    st.close()
