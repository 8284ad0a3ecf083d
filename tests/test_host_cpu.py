"""CPU-only tests: the C-ABI library loads and exports every declared symbol, host loaders/logic follow the
reference, the oracle reproduces the committed golden vectors, multi-rank sharding (gloo, world_size 2)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import _oracle as orc
from genefuserust_b200 import ReadBatch, synth
from genefuserust_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    return _abi.load_library()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "genefuse_gpu.h")).read()
    declared = set(re.findall(r"\b(gf_[a-z_]+)\s*\(", hdr))
    inline = set(re.findall(r"static inline [A-Z_ ]*\w+ (gf_[a-z_]+)\s*\(", hdr))   # defined in the header itself
    assert inline == {"gf_match_order_key"}
    declared -= inline
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for name in declared:
        assert getattr(built, name) is not None
    assert built.gf_abi_version() == 2
    p = _abi.gf_params()
    built.gf_default_params(C.byref(p))
    assert (p.skip_key_dup_threshold, p.major_gene_key_requirement, p.minor_gene_key_requirement,
            p.mismatch_threshold, p.deletion_threshold) == (5, 40, 20, 10, 50)   # src/aux/global_settings.rs:15-29


def test_struct_sizes_match_header():
    assert C.sizeof(_abi.gf_match) == 48
    assert C.sizeof(_abi.gf_batch) == 80
    assert C.sizeof(_abi.gf_params) == 20
    assert C.sizeof(_abi.gf_map_stats) == 120


def test_no_device_fails_loudly(built):
    """no CPU fallback: without a GPU every compute entry point must fail with GF_E_CUDA"""
    if built.gf_device_count() > 0:
        pytest.skip("a GPU is present")
    from genefuserust_b200.host import FusionMapper, GeneFuseError
    with pytest.raises(GeneFuseError) as e:
        FusionMapper.from_gene_spans([(b"ACGT" * 20, False)])
    assert e.value.code == _abi.GF_E_CUDA


def test_product_never_touches_oracle():
    """the product package must not import / link / call anything under oracle/"""
    pkg = os.path.join(ROOT, "genefuserust_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "gf_oracle" not in src and "_oracle" not in src and "orc_" not in src, os.path.join(dp, f)
    out = os.popen(f"ldd {os.path.join(pkg, 'libgenefuse_b200.so')}").read()
    assert "oracle" not in out


def test_parse_csv_and_gene_direction():
    from genefuserust_b200.host import Fusion
    fus = Fusion.parse_csv(os.path.join(GOLD, "testdata", "fusions.csv"))
    assert [f.gene.name for f in fus] == ["ALK", "ROS1", "RET", "EML4"]
    alk = fus[0].gene
    assert (alk.chr, alk.start, alk.end) == ("chr2", 29415640, 30144432)
    assert alk.is_reversed() and fus[1].gene.is_reversed()
    assert not fus[2].gene.is_reversed() and not fus[3].gene.is_reversed()


def test_fasta_reader_tinyref():
    # src/core/fasta_reader.rs:232-279
    from genefuserust_b200.host import FastaReader
    kat = json.load(open(os.path.join(GOLD, "kat.json")))["tinyref"]["strings"]
    ref = FastaReader(os.path.join(GOLD, "testdata", "tinyref.fa")).read_all()
    assert list(ref.m_all_contigs) == ["contig1", "contig2"]
    assert ref.m_all_contigs["contig1"].decode() == kat[1]
    assert ref.m_all_contigs["contig2"].decode() == kat[2]


def test_resolve_gene_spans_rules():
    from genefuserust_b200.host import FastaReader, Fusion, Gene, resolve_gene_spans
    ref = FastaReader("x")
    ref.m_all_contigs = {"chr1": b"acgtACGTNNacgtACGT" * 10, "2": b"TTTTGGGGCCCCAAAA" * 10}
    def fu(name, chrom, a, b, exons):
        g = Gene(name, chrom, a, b)
        for e in exons:
            g.add_exon(*e)
        return Fusion(g)
    fusions = [fu("A", "chr1", 2, 40, [(1, 30, 35), (2, 5, 10)]), fu("B", "1", 0 + 1, 20, []),
               fu("C", "chr2", 4, 36, [(1, 5, 6), (2, 10, 12)]), fu("D", "chrX", 1, 10, [])]
    spans = resolve_gene_spans(ref, fusions)
    assert spans[0] == (ref.m_all_contigs["chr1"][2:40].upper(), True)     # exact name, reversed gene
    assert spans[1] == (ref.m_all_contigs["chr1"][1:20].upper(), False)    # "chr" + name
    assert spans[2] == (ref.m_all_contigs["2"][4:36], False)               # name without "chr"
    assert spans[3] == (b"", False)                                        # unresolved keeps its contig id


def test_testdata_config1_on_oracle():
    """BASELINE config 1 through the oracle: nothing resolves -> empty index -> zero matches"""
    from genefuserust_b200.host import FastaReader, FastqReaderPair, Fusion, resolve_gene_spans
    td = os.path.join(GOLD, "testdata")
    spans = resolve_gene_spans(FastaReader(os.path.join(td, "tinyref.fa")).read_all(),
                               Fusion.parse_csv(os.path.join(td, "fusions.csv")))
    assert spans == [(b"", True), (b"", True), (b"", False), (b"", False)]
    _, batch = FastqReaderPair(os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")).read_all()
    assert batch.n == 3
    idx = orc.OracleIndex(spans)
    assert idx.counts()["n_keys"] == 0
    assert idx.scan(batch) == []
    assert idx.counters()["n_merged"] == 3


def test_oracle_golden_synthetic_matches():
    """committed golden records of a seeded synthetic case (tests/golden/make_golden_matches.py)"""
    g = json.load(open(os.path.join(GOLD, "synth_small_matches.json")))
    panel = synth.make_panel(scale=g["panel_scale"], max_genes=g["max_genes"])
    batch = synth.generate_pairs(panel, g["n_pairs"], read_len=g["read_len"], seed=g["seed"], p_fusion=g["p_fusion"],
                                 threads=2)
    import hashlib
    assert hashlib.sha256(batch.seq1.tobytes() + batch.seq2.tobytes()).hexdigest() == g["reads_sha256"]
    idx = orc.OracleIndex(panel.genes())
    assert idx.counts() == g["index_counts"]
    got = idx.scan(batch, threads=4)
    assert [list(r) for r in got] == g["matches"]


def test_oracle_segment_mask_properties():
    import random
    rng = random.Random(1)
    for _ in range(300):
        n = rng.randint(2, 80)
        mask = bytes(rng.choice((0, 1, 2, 2, 3, 3, 3)) for _ in range(n))
        segs = orc.segment_mask(mask, (1 << 32) | 5, (2 << 32) | 9)
        for s, e, c, p in segs:
            assert e - s > 20 and mask[s] in (2, 3) and mask[e] == mask[s] and s < n - 1
            tgt = mask[s]
            assert all(m <= tgt for m in mask[s:e + 1])
    # the last index can never open a segment; SECOND runs are cut by TOP
    assert orc.segment_mask(bytes([3] * 30), 1 << 32, 2 << 32) == [(0, 29, 1, 0)]
    assert orc.segment_mask(bytes([2] * 25 + [3] + [2] * 25), 1 << 32, 2 << 32) == [(0, 24, 2, 0)]
    assert orc.segment_mask(bytes([3] * 15 + [0] * 9 + [3] * 15), 1 << 32, 2 << 32) == [(0, 38, 1, 0)]
    assert orc.segment_mask(bytes([3] * 15 + [0] * 10 + [3] * 15), 1 << 32, 2 << 32) == []


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genefuserust_b200 import sharding
    panel = synth.make_panel(scale=0.01, max_genes=24)
    batch = synth.generate_pairs(panel, 6001, read_len=150, seed=3, p_fusion=0.05, threads=1)
    idx = orc.OracleIndex(panel.genes())
    local = sharding.map_shard(lambda b: idx.scan(b, threads=1), batch, rank, world)
    allrec = sharding.gather_matches(local)
    if rank == 0:
        q.put((allrec, idx.scan(batch, threads=2)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_world2_gloo():
    """N>1 host logic on CPU: each rank maps its shard (oracle stands in for the device), records are gathered;
    the union must equal the unsharded result."""
    import torch.multiprocessing as mp
    from genefuserust_b200 import sharding
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(want) > 5 and got == want


def test_add_match_bucket_and_sort_order():
    """bucket = n_genes * right.contig + left.contig (fusion_mapper.rs:263); sort_matches orders a
    bucket by read_break desc, read length asc, read name desc, stable (read_match.rs:203-229, fusion_mapper.rs:379-385)"""
    from genefuserust_b200 import host
    from genefuserust_b200._abi import gf_match

    def mk(pair, rb, slen, lc=1, rc=2):
        m = gf_match()
        m.pair_idx, m.read_break, m.seq_len, m.l_contig, m.r_contig = pair, rb, slen, lc, rc
        return m
    assert host.fusion_bucket(136, mk(0, 10, 100, lc=3, rc=7)) == 136 * 7 + 3
    names = {0: "r0", 1: "r1", 2: "r2", 3: "r3", 4: "r4", 5: "r5", 6: "r5"}
    ms = [mk(0, 50, 150), mk(1, 70, 150), mk(2, 70, 120), mk(3, 70, 120), mk(4, 50, 150), mk(5, 60, 99), mk(6, 60, 99)]
    host.sort_read_matches(ms, lambda m: names[m.pair_idx])
    # break 70 first (len 120 before 150; among the two 120s the larger name first), then 60 (equal names: push order kept),
    # then 50 (r4 before r0)
    assert [m.pair_idx for m in ms] == [3, 2, 1, 5, 6, 4, 0]


def test_header_is_plain_c(tmp_path):
    """include/genefuse_gpu.h is the drop-in boundary: it must compile as C99 (plain pointers and sizes, no C++) and agree
    with the ctypes mirror on the struct sizes"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "genefuse_gpu.h"\n#include <stdio.h>\nint main(void){ gf_match m = {0}; '
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %llu\\n", '
                   'sizeof(gf_match), sizeof(gf_batch), sizeof(gf_params), sizeof(gf_map_stats), sizeof(gf_break_ref), '
                   'sizeof(gf_break_job), sizeof(gf_break_out), sizeof(gf_ref_contig), sizeof(gf_reference_info), '
                   'sizeof(gf_alignable_result), (unsigned long long)gf_match_order_key(136, &m)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_abi.gf_match), C.sizeof(_abi.gf_batch), C.sizeof(_abi.gf_params), C.sizeof(_abi.gf_map_stats),
                     C.sizeof(_abi.gf_break_ref), C.sizeof(_abi.gf_break_job), C.sizeof(_abi.gf_break_out),
                     C.sizeof(_abi.gf_ref_contig), C.sizeof(_abi.gf_reference_info), C.sizeof(_abi.gf_alignable_result),
                     _abi.gf_match_order_key(136, _abi.gf_match())]
