#!/usr/bin/env python3
"""developer tool: call cost of the C ABI at small batch sizes, measured by the plain-C driver tests/c_driver/replay.c
(per-call gf_map_pairs at 1 k / 64 k / 1 M pairs vs the batched shim fed with 1000-pair packs), full cancer-shaped panel.
usage (under gpurun): python tools/small_batch.py [pairs] > gpurun_out/small_batch.jsonl"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from genefuserust_b200 import synth
import test_c_driver as T

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
with tempfile.TemporaryDirectory() as d:
    exe = T.build_driver(d)
    panel = synth.make_panel()
    b = synth.generate_pairs(panel, P, read_len=150, seed=12, threads=16)
    dump = os.path.join(d, "run.bin")
    T.write_dump(dump, panel.genes(), b, 150)
    sys.stdout.write(subprocess.check_output([exe, dump]).decode())
