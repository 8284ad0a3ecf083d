"""GPU debug: categorise the screen's survivors (kind of read, merged?, true gate passers per the oracle).
Developer tool (test infrastructure): the oracle is used only as the checker."""
import collections, ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as ge
ge.build()
from genefuserust_b200 import synth, ReadBatch
from genefuserust_b200.host import FusionMapper
import _oracle
L = int(os.environ.get("L", "150"))
N = int(os.environ.get("N", "300000"))
panel = synth.make_panel(scale=float(os.environ.get("SCALE", "1.0")))
batch, kind = synth.generate_pairs(panel, N, read_len=L, seed=13, return_kind=True)
m = FusionMapper.from_gene_spans(panel.genes(), device=0)
got = m.scan_pair_end(batch)
st = m.map_stats()
print("L", L, "pairs", N, "survivors", st.n_survivors, "matches", len(got), "seqs", st.n_sequences)
buf = np.zeros((4 * N, 2), dtype=np.uint32)
n = C.c_uint64(0)
m.lib.gf_debug_get_survivors(C.c_void_p(m.m_indexer.h.value), C.c_void_p(buf.ctypes.data), C.c_uint64(4 * N), C.byref(n))
print("NOTE: survivor list only covers the last chunk of the host pipeline:", n.value)
sv = buf[:min(n.value, 4 * N)]
o = _oracle.OracleIndex(panel.genes())
cat = collections.Counter()
ex = {}
for pair, meta in sv[:3000]:
    src = int(meta & 3)
    # absolute pair index unknown for chunked runs -> debug runs use a single chunk (small N)
    s1, q1 = batch.read(int(pair), 1); s2, q2 = batch.read(int(pair), 2)
    if src == 0:
        seq = _oracle.fast_merge(s1, q1, s2, q2)[0]
    else:
        seq = s1 if src == 1 else s2
    segs = o.map_read(seq)
    key = (int(kind[pair]), src, len(segs))
    cat[key] += 1
    ex.setdefault(key, (int(pair), src, len(seq)))
print("(kind 0=target 1=off 2=fusion, source, oracle segments) -> count   [first 3000 survivors]")
for k, v in sorted(cat.items()):
    print(k, v, "e.g.", ex[k])
