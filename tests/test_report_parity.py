"""Report-level parity (BASELINE.json north_star: "the same unique read counts per fusion, byte-identical fusion reports").

Two pipelines over the same synthetic run (testdata/fusions.csv rebased onto synthetic contigs, reads with planted fusions):
  A  all CPU: oracle records -> per-record filters -> Matcher (oracle) -> sort_matches -> cluster_matches with the oracle's
     adjust_fusion_break -> JSON report                                  (tests/report_port.py restates the host stages)
  B  device results feeding the same host stage: gf_map_pairs with GF_OUT_DROP_FILTERED | GF_OUT_BUCKET_ORDER,
     gf_alignable_filter, gf_adjust_fusion_break -> the same cluster / report code
and the JSON bytes must be identical.  CPU-only tests pin the restated host stages to what the reference offers: the pos2str
vectors of src/core/fusion.rs:116-141 and the derivable report of BASELINE config 1.
"""
import json
import os

import numpy as np
import pytest

import _oracle as orc
import report_port as rp
from genefuserust_b200 import ReadBatch, host, synth
from genefuserust_b200._abi import GF_OUT_BUCKET_ORDER, GF_OUT_DROP_FILTERED, gf_match

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def test_pos2str_vectors_of_the_reference():
    kat = json.load(open(os.path.join(GOLD, "kat.json")))["pos2str"]
    genes = {g.m_name: g for g in rp.parse_csv(os.path.join(GOLD, "testdata", "fusions.csv"))}
    assert set(genes) == {"ALK", "ROS1", "RET", "EML4"}
    for name, pos, want in kat["cases"]:
        assert genes[name].pos2str(pos) == want
    # the host mirror used by the product-side tests parses the same file the same way
    mirror = {f.gene.name: f.gene for f in host.Fusion.parse_csv(os.path.join(GOLD, "testdata", "fusions.csv"))}
    for name, g in genes.items():
        assert (mirror[name].chr, mirror[name].start, mirror[name].end, mirror[name].reversed) == (g.m_chr, g.m_start, g.m_end, g.m_reversed)


def test_config1_report_is_the_derivable_one():
    """BASELINE config 1: tinyref.fa holds none of chr2 / chr6 / chr10, every gene is empty, nothing matches:
    "fusions":{ } with the exact bytes of json_reporter.rs:37-41,109"""
    body = rp.json_report([], command="genefuse -r tinyref.fa -f fusions.csv -1 R1.fq -2 R2.fq", time_str="T")
    assert body == (b'{\n\t"command":"genefuse -r tinyref.fa -f fusions.csv -1 R1.fq -2 R2.fq",\n\t"version":"0.1.2",\n'
                    b'\t"time":"T",\n\t"fusions":{\n\t}\n}\n\n')
    ref = host.FastaReader(os.path.join(GOLD, "testdata", "tinyref.fa")).read_all()
    spans = host.resolve_gene_spans(ref, host.Fusion.parse_csv(os.path.join(GOLD, "testdata", "fusions.csv")))
    assert [s for s, _ in spans] == [b"", b"", b"", b""]


# ------------------------------------------------------------------------------------------------ the synthetic run
def build_case(tmp_path, n_pairs=120_000, seed=31):
    """fusions.csv rebased onto synthetic contigs (gene order, names, exon layout and strands kept), a FASTA with N
    gaps in the spacers (like a real assembly: every base code starts many ACGT runs, which keeps the reference's Matcher off
    its panic path), and read pairs with planted fusions between the four genes"""
    genes = rp.parse_csv(os.path.join(GOLD, "testdata", "fusions.csv"))
    by_chr = {}
    for g in genes:
        by_chr.setdefault(g.m_chr, []).append(g)
    rng = np.random.RandomState(seed)
    contigs = {}
    csv_lines = []
    new_genes = {}
    for ci, (chrom, gl) in enumerate(sorted(by_chr.items())):
        cur = 10_000
        layout = []
        for g in gl:
            delta = cur - g.m_start
            layout.append((g, delta))
            cur += (g.m_end - g.m_start) + 10_000
        seq = synth.random_bases(seed * 131 + ci, cur).copy()
        # N gaps, only inside the spacers
        for g, delta in layout:
            s0 = g.m_start + delta - 10_000
            for k in range(150):
                p = s0 + 20 + 60 * k
                seq[p:p + int(rng.randint(1, 4))] = ord("N")
        contigs[chrom] = seq
        for g, delta in layout:
            new_genes[g.m_name] = (g, delta)
    for g in genes:                    # CSV order = contig ids
        g0, delta = new_genes[g.m_name]
        csv_lines.append(f">{g.m_name}_ENST,{g.m_chr}:{g.m_start + delta}-{g.m_end + delta}")
        for e in g.m_exons:
            csv_lines.append(f"{e.id},{e.start + delta},{e.end + delta}")
    csv_path = os.path.join(tmp_path, "fusions.synth.csv")
    open(csv_path, "w").write("\n".join(csv_lines) + "\n")
    fa_path = os.path.join(tmp_path, "ref.synth.fa")
    with open(fa_path, "wb") as f:
        for chrom, seq in contigs.items():
            f.write(b">" + chrom.encode() + b"\n")
            b = seq.tobytes()
            f.write(b"\n".join(b[i:i + 60] for i in range(0, len(b), 60)) + b"\n")
    ref = host.FastaReader(fa_path).read_all()
    fusions = host.Fusion.parse_csv(csv_path)
    spans = host.resolve_gene_spans(ref, fusions)
    rgenes = rp.parse_csv(csv_path)
    assert [len(s) for s, _ in spans] == [g.m_end - g.m_start for g in rgenes] and all(len(s) for s, _ in spans)
    # planted fusions: (gene_a, pos_a, strand_a, gene_b, pos_b, strand_b), all four strand combinations
    ng = len(spans)
    planted = []
    for k in range(8):
        ga, gb = k % ng, (k + 1 + k // ng) % ng
        la, lb = len(spans[ga][0]), len(spans[gb][0])
        planted.append((ga, int(rng.randint(la // 4, 3 * la // 4)), 1 if (k & 1) == 0 else -1,
                        gb, int(rng.randint(lb // 4, 3 * lb // 4)), 1 if (k & 2) == 0 else -1))
    panel = synth.Panel([g.m_name for g in rgenes], [np.frombuffer(s, dtype=np.uint8).copy() for s, _ in spans],
                        [int(r) for _, r in spans], planted)
    batch = synth.generate_pairs(panel, n_pairs, read_len=150, seed=seed, p_fusion=0.03, threads=8)
    n1 = [b"@SYN:%d:%d 1:N:0:ACGT" % (seed, i // 3) for i in range(batch.n)]      # groups of 3 pairs share a name: ties
    n2 = [b"@SYN:%d:%d 2:N:0:ACGT" % (seed, i // 3) for i in range(batch.n)]
    return {"ref": ref, "spans": spans, "rgenes": rgenes, "batch": batch, "names": (n1, n2),
            "contigs": [ref.m_all_contigs[k] for k in sorted(ref.m_all_contigs)]}


def to_struct(t):
    r = gf_match()
    for f, v in zip(gf_match.FIELDS, t):
        setattr(r, f, v)
    return r


def bucketize(rms, n_genes):
    fm = {}
    for rm in rms:
        fm.setdefault(n_genes * rm.m_right_gp.contig + rm.m_left_gp.contig, []).append(rm)       # add_match, :263
    return fm


def oracle_adjust(batch_of_results):
    out = []
    for lref, rref, matches in batch_of_results:
        res = []
        for seq, rb in matches:
            shift, ld, rd, status = orc.adjust_fusion_break(seq, rb, lref, rref)
            assert status == 0
            res.append((shift, ld, rd))
        out.append(res)
    return out


def pipeline_cpu(case):
    """the reference's flow on the CPU: scan (oracle) -> filter_matches -> sort_matches -> cluster_matches -> JSON"""
    spans, batch, names = case["spans"], case["batch"], case["names"]
    o = orc.OracleIndex(spans)
    recs = [to_struct(t) for t in o.scan(batch, threads=8)]
    o.close()
    rms = [rp.make_read_match(r, rp.read_of_record(r, batch, names, orc.fast_merge)) for r in recs]
    # the per-record filters, from the restated predicates; the records' own flags (oracle scan) must say the same
    for rm in rms:
        assert rp.per_record_filter_flags(rm) == rm.filter_flags
    n_before = len(rms)
    rms = [rm for rm in rms if rm.filter_flags == 0]
    fm = bucketize(rms, len(spans))
    order = [rm for b in sorted(fm) for rm in fm[b]]                      # remove_alignables' gathering order (:496-500)
    flags, res, rc = orc.remove_alignables(case["contigs"], [rm.m_read.m_seq for rm in order])
    assert rc == 0 and not flags.any(), (rc, res.astuple())               # nothing removed, no panic
    rp.sort_matches(fm)
    results = rp.cluster_matches(fm, len(spans) ** 2, case["rgenes"], [s for s, _ in spans], orc.edit_distance, oracle_adjust)
    return rp.json_report(results), results, (n_before, len(rms), res.astuple())


def pipeline_gpu(case):
    """device results feeding the same host stage"""
    spans, batch, names = case["spans"], case["batch"], case["names"]
    m = host.FusionMapper.from_gene_spans(spans, device=0)
    m.set_output_mode(GF_OUT_DROP_FILTERED | GF_OUT_BUCKET_ORDER)
    recs = m.scan_pair_end(batch)
    n_before = int(m.map_stats().n_matches)
    reads = {(r.pair_idx, r.source): rp.read_of_record(r, batch, names, orc.fast_merge) for r in recs}
    recs = m.finish_order(recs, lambda r: reads[(r.pair_idx, r.source)].m_name)
    rms = [rp.make_read_match(r, reads[(r.pair_idx, r.source)]) for r in recs]
    fm = {}
    for rm in rms:                       # already in bucket order, every bucket in sort_matches order
        fm.setdefault(len(spans) * rm.m_right_gp.contig + rm.m_left_gp.contig, []).append(rm)
    # remove_alignables runs BEFORE sort_matches in the reference (fusion_mapper.rs:291-295): its retain() visits the buckets
    # in push order; the result (nothing removed / which sequence panics first) is reported for that order
    push = [rm for b in sorted(fm) for rm in sorted(fm[b], key=lambda x: x.push_key)]
    mt = host.Matcher(case["contigs"], device=0)
    flags, res, rc = mt.remove_alignables([rm.m_read.m_seq for rm in push])
    mt.close()
    assert rc == 0 and not flags.any(), (rc, res.astuple())

    def gpu_adjust(batch_of_results):
        out = m.adjust_fusion_break(batch_of_results)
        return [[(s, ld, rd) for (s, ld, rd, status) in r] for r in out]
    results = rp.cluster_matches(fm, len(spans) ** 2, case["rgenes"], [s for s, _ in spans], orc.edit_distance, gpu_adjust)
    body = rp.json_report(results)
    m.close()
    return body, results, (n_before, len(rms), res.astuple())


def test_cpu_pipeline_reports_planted_fusions(tmp_path):
    case = build_case(str(tmp_path), n_pairs=60_000)
    body, results, (n_before, n_after, matcher) = pipeline_cpu(case)
    assert n_before > n_after > 100
    assert len(results) >= 3, [r.m_title for r in results]
    assert all(r.m_unique >= 2 for r in results)
    titles = [r.m_title for r in results]
    assert all(t.startswith("Fusion: ") for t in titles)
    assert body.count(b'"unique":') == len(results) and body.endswith(b"\n\t}\n}\n\n")
    # sort_fusion_results: descending by unique reads
    assert [r.m_unique for r in results] == sorted((r.m_unique for r in results), reverse=True)
    assert min(matcher[0]) > 50          # every base code starts more than 50 runs: the Matcher casts no votes


@pytest.mark.gpu
def test_report_bytes_identical_with_device_stages(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    case = build_case(str(tmp_path))
    body_cpu, res_cpu, info_cpu = pipeline_cpu(case)
    body_gpu, res_gpu, info_gpu = pipeline_gpu(case)
    assert info_cpu == info_gpu, (info_cpu, info_gpu)
    assert [(r.m_title, r.m_unique, len(r.m_matches)) for r in res_cpu] == [(r.m_title, r.m_unique, len(r.m_matches)) for r in res_gpu]
    assert body_cpu == body_gpu
    assert len(res_cpu) >= 4 and len(body_cpu) > 10_000
