"""GPU debug: categorise the screen's survivors (kind of read, merged?, length parity, strand)."""
import collections, ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as ge
ge.build()
from genefuserust_b200 import synth
from genefuserust_b200.host import FusionMapper
import _oracle
panel = synth.make_panel(scale=float(os.environ.get("SCALE", "0.05")))
N = 200000
batch, kind = synth.generate_pairs(panel, N, read_len=150, seed=12, return_kind=True)
m = FusionMapper.from_gene_spans(panel.genes(), device=0)
# single chunk so that the survivor list covers the whole batch
got = m.scan_pair_end(batch)
st = m.map_stats()
print("pairs", N, "survivors", st.n_survivors, "matches", len(got), "seqs", st.n_sequences)
buf = np.zeros((4 * N, 2), dtype=np.uint32)
n = C.c_uint64(0)
m.lib.gf_debug_get_survivors(C.c_void_p(m.m_indexer.h.value), C.c_void_p(buf.ctypes.data), C.c_uint64(4 * N), C.byref(n))
sv = buf[:n.value]
cat = collections.Counter()
ex = {}
for pair, meta in sv:
    src = meta & 3
    olen = (meta >> 2) & 0xFFF
    if src == 0:
        s1, q1 = batch.read(int(pair), 1); s2, q2 = batch.read(int(pair), 2)
        ms = _oracle.fast_merge(s1, q1, s2, q2)
        ln = len(ms[0])
    else:
        ln = 150
    key = (int(kind[pair]), int(src), ln & 1)
    cat[key] += 1
    ex.setdefault(key, (int(pair), int(src), ln))
print("(kind 0=target 1=off 2=fusion, source 0=merged 1=R1 2=R2, len&1) -> count")
for k, v in sorted(cat.items()):
    print(k, v, "e.g.", ex[k])
tot = collections.Counter(int(k) for k in kind)
print("kinds of pairs", tot)
