"""small end-to-end run for compute-sanitizer: PE + SE, ragged + long reads, FASTQ text / streams (plain + gzip), packs, output
modes, the device entry points in several chunks, adjust_fusion_break, the Matcher pass (developer tool)"""
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
# output modes, pack stream, FASTQ streams, chunked device path, Matcher pass
import ctypes as C, gzip, torch
from genefuserust_b200 import host
from genefuserust_b200._abi import gf_batch, gf_match
b = synth.generate_pairs(panel, 4000, read_len=150, seed=7, p_fusion=0.2, threads=2)
m.set_output_mode(3)
print("mode 3", len(m.scan_pair_end(b)))
m.set_output_mode(0)
st = host.PackStream(m, paired=True, batch_pairs=1500)
for lo in range(0, b.n, 1000):
    st.push(lo, [b.read(i, 1) for i in range(lo, min(lo + 1000, b.n))], [b.read(i, 2) for i in range(lo, min(lo + 1000, b.n))])
st.flush()
print("packs", len(st.take()))
st.close()
f1, f2 = fq(b.seq1, b.qual1, b.off1), fq(b.seq2, b.qual2, b.off2)
for gz in (False, True):
    fs = host.FastqStream(m, paired=True, gz=gz, chunk_bytes=1 << 16)
    e1, e2 = (gzip.compress(f1), gzip.compress(f2)) if gz else (f1, f2)
    for p in range(0, max(len(e1), len(e2)), 50_001):
        fs.feed(e1[p:p + 50_001], e2[p:p + 50_001])
    fs.finish()
    print("fastq stream gz" if gz else "fastq stream", len(fs.take()), fs.counts())
    fs.close()
os.environ["GF_DEVICE_CHUNK_PAIRS"] = "1500"
d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (b.seq1, b.qual1, b.seq2, b.qual2)]
o1 = torch.from_numpy(b.off1.view(np.int64)).cuda()
db = gf_batch()
db.n = b.n
db.seq1, db.qual1, db.seq2, db.qual2 = (t.data_ptr() for t in d)
db.off1 = db.off2 = o1.data_ptr()
db.bytes1 = db.bytes2 = int(b.off1[-1])
db.max_len = 150
d_out = torch.empty(2 * b.n * C.sizeof(gf_match), dtype=torch.uint8, device="cuda")
d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
assert m.lib.gf_map_pairs_device(m.m_indexer.h, C.byref(db), d_out.data_ptr(), 2 * b.n, d_n.data_ptr(), None) == 0
torch.cuda.synchronize()
print("device path, chunked", int(d_n.item()))
rng = np.random.default_rng(1)
contigs = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy() for n in (300_000, 70_001, 17, 16)]
contigs[0][1000:1100] = ord("N")
contigs[0][5000:5040] = ord("A")
mt = host.Matcher(contigs)
print("matcher", mt.remove_alignables([b"ACGT" * 20, b"acgtn" * 10])[1].astuple(), mt.info().n_bases)
mt.close()
dc = [torch.from_numpy(c).cuda() for c in contigs]
mt = host.Matcher([(t.data_ptr(), t.numel()) for t in dc])
print("matcher (device contigs)", mt.remove_alignables([b"ACGT" * 20])[1].astuple())
mt.close()
# packed / hybrid upload from pinned arenas (ragged reads with N and lower case), several chunks
rng = random.Random(3)
r1 = [(bytes(rng.choice(b"ACGTACGTACGTNacgt") for _ in range(n)), b"E" * n) for n in [0, 1, 31, 32, 33, 64, 65, 100, 150, 151, 200, 256] * 300]
r2 = [(bytes(rng.choice(b"ACGTACGTACGTNacgt") for _ in range(n)), b"E" * n) for n in [256, 200, 151, 150, 100, 65, 64, 33, 32, 31, 1, 0] * 300]
b = ReadBatch.from_reads(r1, r2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
bp = ReadBatch(pin(b.seq1), pin(b.qual1), pin(b.off1.view(np.int64)).view(np.uint64), pin(b.seq2), pin(b.qual2), pin(b.off2.view(np.int64)).view(np.uint64))
bp.max_len = b.max_len
os.environ["GF_CHUNK_MB"] = "1"
for mode in ("0", "1", "2"):
    os.environ["GF_HOST_PACK"] = mode
    print("upload mode", mode, len(m.scan_pair_end(bp)), m.map_stats().packed_upload)
del os.environ["GF_HOST_PACK"], os.environ["GF_CHUNK_MB"]
m.close()
print("ok")
