"""small end-to-end run for compute-sanitizer (all screen versions, PE + SE, ragged + long reads)"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
ge.build()
from genefuserust_b200 import synth, ReadBatch
from genefuserust_b200.host import FusionMapper
panel = synth.make_panel(scale=0.01, max_genes=24)
m = FusionMapper.from_gene_spans(panel.genes(), device=0)
for L in (75, 150, 250):
    b = synth.generate_pairs(panel, 3000, read_len=L, seed=5, p_fusion=0.2, threads=2)
    print(L, len(m.scan_pair_end(b)), len(m.scan_single_end(ReadBatch(b.seq1, b.qual1, b.off1))), len(m.fast_merge(b)))
rng = random.Random(1)
r1 = [(bytes(rng.choice(b"ACGTN") for _ in range(n)), b"E" * n) for n in [0, 1, 15, 16, 31, 33, 100, 257, 640, 1000] * 20]
r2 = [(bytes(rng.choice(b"ACGTN") for _ in range(n)), b"E" * n) for n in [1000, 640, 257, 100, 33, 31, 16, 15, 1, 0] * 20]
b = ReadBatch.from_reads(r1, r2)
print("ragged", len(m.scan_pair_end(b)))
# raw FASTQ text through the device-side ingest, and the report-stage break adjustment
b = synth.generate_pairs(panel, 500, read_len=100, seed=6, p_fusion=0.3, threads=2)
fq = lambda seq, qual, off: b"".join(b"@r%d\n" % i + bytes(seq[off[i]:off[i + 1]]) + b"\n+\n" + bytes(qual[off[i]:off[i + 1]]) + b"\n" for i in range(b.n))
print("fastq", len(m.scan_fastq(fq(b.seq1, b.qual1, b.off1), fq(b.seq2, b.qual2, b.off2))[0]))
g0, g1 = panel.genes()[0][0], panel.genes()[1][0]
read = g0[100:160] + g1[200:260]
print("adjust", m.adjust_fusion_break([(g0[80:160], g1[200:280], [(read, 59), (read, 57), (read, 62), (read, 1)]), (b"", b"", [(read, 40)])]))
m.close()
print("ok")
