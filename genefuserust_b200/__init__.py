"""genefuserust_b200 — B200-native (sm_100a CUDA behind a C ABI) implementation of
GeneFuseRust's per-read fusion-matching hot path.  See DESIGN.md."""
from ._abi import (gf_batch, gf_gene_span, gf_index_info, gf_lookup, gf_map_stats, gf_match, gf_merge_info,
                   gf_params, load_library)
from .batch import ReadBatch

__all__ = ["gf_batch", "gf_gene_span", "gf_index_info", "gf_lookup", "gf_map_stats", "gf_match",
           "gf_merge_info", "gf_params", "load_library", "ReadBatch"]
