"""Randomised parity fuzz on the GPU: many small panels / read lengths / error rates / fusion rates, the CUDA path
(PE, SE, list mode) against the CPU oracle.  usage: python tools/fuzz_parity.py [iterations] [seed]

Developer tool (test / measurement infrastructure, not product code): the CPU oracle is loaded here only as the checker of
the CUDA path's records and as the reported CPU rate; the package under genefuserust_b200/ never touches it."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
ge.build()
from genefuserust_b200 import synth, ReadBatch
from genefuserust_b200 import host
import _oracle as orc

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
n_pairs = n_matches = 0
for it in range(iters):
    scale = rng.choice((0.004, 0.01, 0.02, 0.05))
    max_genes = rng.choice((None, 8, 24, 60))
    panel = synth.make_panel(seed=20240201 + it, scale=scale, max_genes=max_genes, n_fusions=rng.choice((4, 20, 60)))
    L = rng.choice((36, 50, 75, 100, 125, 150, 151, 160, 161, 200, 250, 256, 257, 300))
    n = rng.choice((2000, 20000, 60000))
    kw = dict(read_len=L, seed=rng.randrange(1 << 30), p_target=rng.choice((0.2, 0.5, 0.8)), p_fusion=rng.choice((0.001, 0.05, 0.3, 0.7)),
              sub_rate=rng.choice((0.0, 0.002, 0.01, 0.03)), n_rate=rng.choice((0.0, 0.0005, 0.005)))
    b = synth.generate_pairs(panel, n, **kw)
    if it % 3 == 2:
        # ragged lengths + lower case / IUPAC / N characters (the converters' slow paths, every alignment class)
        n = min(n, 20000)
        alpha = b"acgtnNRYKM."
        def mutate(seq, qual, off, i):
            a, e = int(off[i]), int(off[i + 1])
            cut = rng.randint(0, min(40, e - a)) if rng.random() < 0.7 else 0
            s_ = bytearray(seq[a:e - cut].tobytes())
            for _ in range(rng.choice((0, 0, 1, 2, 5))):
                if s_:
                    s_[rng.randrange(len(s_))] = rng.choice(alpha)
            return bytes(s_), qual[a:e - cut].tobytes()
        r1 = [mutate(b.seq1, b.qual1, b.off1, i) for i in range(n)]
        r2 = [mutate(b.seq2, b.qual2, b.off2, i) for i in range(n)]
        b = ReadBatch.from_reads(r1, r2)
    genes = panel.genes()
    m = host.FusionMapper.from_gene_spans(genes, device=0)
    o = orc.OracleIndex(genes)
    want = o.scan(b, threads=os.cpu_count() or 8)
    got = [r.astuple() for r in m.scan_pair_end(b)]
    assert got == want, ("PE", it, scale, max_genes, L, n, kw, len(got), len(want))
    if b.max_len <= 256:
        # the same batch from pinned arenas: packed upload of every chunk (GF_HOST_PACK=1), then the per-chunk mix (default),
        # with small pipeline chunks
        import torch
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        bp = ReadBatch(pin(b.seq1), pin(b.qual1), pin(b.off1.view(np.int64)).view(np.uint64),
                       pin(b.seq2), pin(b.qual2), pin(b.off2.view(np.int64)).view(np.uint64))
        bp.max_len = b.max_len
        os.environ["GF_CHUNK_MB"] = "1"
        for mode in ("1", "2"):
            os.environ["GF_HOST_PACK"] = mode
            os.environ["GF_PACK_THREADS"] = str(rng.choice((8, 11, 16)))
            got_p = [r.astuple() for r in m.scan_pair_end(bp)]
            assert got_p == want, ("PE packed upload", mode, it, scale, max_genes, L, n, kw, len(got_p), len(want))
        for k_ in ("GF_CHUNK_MB", "GF_HOST_PACK", "GF_PACK_THREADS"):
            os.environ.pop(k_, None)
    se = ReadBatch(b.seq2, b.qual2, b.off2)
    assert [r.astuple() for r in m.scan_single_end(se)] == o.scan(se, threads=os.cpu_count() or 8), ("SE", it, L, kw)
    if it % 5 == 0:
        sub = genes[: max(2, len(genes) // 2)]
        m2 = host.FusionMapper.from_gene_spans(sub, device=0)
        lst = host.scan_list([m, m2], b)
        o2 = orc.OracleIndex(sub)
        assert [r.astuple() for r in lst[0]] == want and [r.astuple() for r in lst[1]] == o2.scan(b, threads=os.cpu_count() or 8), ("list", it)
        m2.close(); o2.close()
    if it % 4 == 1:
        # device-side record filter + bucket order against the oracle twin of add_match / sort_matches (names with ties)
        from genefuserust_b200._abi import gf_match
        recs = []
        for t in want:
            r = gf_match()
            for f, v in zip(gf_match.FIELDS, t):
                setattr(r, f, v)
            recs.append(r)
        name = lambda r: b"@r%d/%d" % (r.pair_idx // 5, 0 if r.source == 0 else r.source)
        order = orc.bucket_sort(recs, m.n_genes, [name(r) for r in recs], True)
        m.set_output_mode(3)
        got3 = m.finish_order(m.scan_pair_end(b), name)
        m.set_output_mode(0)
        assert [r.astuple() for r in got3] == [want[i] for i, _ in order], ("bucket order", it)
    n_pairs += n; n_matches += len(want)
    print(f"  it {it:3d}: scale {scale} genes {len(genes):3d} L {L:3d} pairs {n:6d} p_fusion {kw['p_fusion']} sub {kw['sub_rate']} n {kw['n_rate']}"
          f"{' ragged+IUPAC' if it % 3 == 2 else ''}{' +list' if it % 5 == 0 else ''}{' +order' if it % 4 == 1 else ''}: {len(want)} records identical (PE{', PE packed / mixed upload' if b.max_len <= 256 else ''}), SE identical", flush=True)
    m.close(); o.close()
print(f"fuzz ok: {iters} configurations, {n_pairs} pairs, {n_matches} matches, {time.time() - t0:.1f} s")
