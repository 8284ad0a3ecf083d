import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
from genefuserust_b200 import synth
from genefuserust_b200.host import FusionMapper
panel = synth.make_panel(scale=1.0)
genes = panel.genes()
for rep in range(3):
    t = time.perf_counter()
    m = FusionMapper.from_gene_spans(genes, device=0)
    print("create", round(time.perf_counter() - t, 3), "s  build_ms", round(m.m_indexer.info().build_ms, 2), flush=True)
    m.close()
