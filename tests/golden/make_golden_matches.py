#!/usr/bin/env python3
"""Writes tests/golden/synth_small_matches.json: the oracle's records on a small seeded synthetic case.  The oracle
cannot be checked against the real reference here (no Rust toolchain), so this fixture pins the oracle against
ITSELF over time (regression) and gives the GPU tests a committed target."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as orc  # noqa: E402
from genefuserust_b200 import synth  # noqa: E402

cfg = {"panel_scale": 0.01, "max_genes": 24, "n_pairs": 8000, "read_len": 150, "seed": 12, "p_fusion": 0.03}
panel = synth.make_panel(scale=cfg["panel_scale"], max_genes=cfg["max_genes"])
batch = synth.generate_pairs(panel, cfg["n_pairs"], read_len=cfg["read_len"], seed=cfg["seed"], p_fusion=cfg["p_fusion"], threads=2)
idx = orc.OracleIndex(panel.genes())
cfg["reads_sha256"] = hashlib.sha256(batch.seq1.tobytes() + batch.seq2.tobytes()).hexdigest()
cfg["index_counts"] = idx.counts()
cfg["matches"] = [list(r) for r in idx.scan(batch, threads=4)]
cfg["fields"] = ["pair_idx", "source", "used_rc", "reversed", "read_break", "l_contig", "l_pos", "r_contig", "r_pos",
                 "gap", "l_dist", "r_dist", "seq_len", "merge_olen", "merge_diff", "filter_flags"]
json.dump(cfg, open(os.path.join(HERE, "synth_small_matches.json"), "w"))
print(len(cfg["matches"]), "matches")
