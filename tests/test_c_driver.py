"""A plain-C program (tests/c_driver/replay.c) links libgenefuse_b200.so and drives the C ABI the way the Rust shim would:
per-call gf_map_pairs at 1 k / 64 k / 1 M pairs and the batched shim (gf_stream_*) fed with 1000-pair packs.  CPU: it compiles
as C99 against the header.  -m gpu: every mode returns the same records (checksum) as the Python path."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "replay.c")
PKG = os.path.join(ROOT, "genefuserust_b200")


def build_driver(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    exe = os.path.join(str(tmp_path), "replay")
    subprocess.check_call(["gcc", "-std=c99", "-D_POSIX_C_SOURCE=199309L", "-O2", "-Wall", "-Wextra", "-Werror", "-I",
                           os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", PKG, "-lgenefuse_b200", "-Wl,-rpath," + PKG])
    return exe


def write_dump(path, genes, batch, L):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(genes)))
        for seq, rev in genes:
            f.write(struct.pack("<IB", len(seq), 1 if rev else 0))
            f.write(seq)
        f.write(struct.pack("<QI", batch.n, L))
        for a in (batch.seq1, batch.qual1, batch.seq2, batch.qual2):
            f.write(np.ascontiguousarray(a).tobytes())


def fnv(recs):
    h = 1469598103934665603
    for b in recs:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_c_driver_compiles_as_c99(tmp_path):
    assert os.path.exists(build_driver(tmp_path))


@pytest.mark.gpu
def test_c_driver_replay_matches_python_path(tmp_path):
    import ctypes as C
    from genefuserust_b200 import host, synth
    from genefuserust_b200._abi import gf_match
    exe = build_driver(tmp_path)
    panel = synth.make_panel(scale=0.02)
    L = 150
    b = synth.generate_pairs(panel, 70_000, read_len=L, seed=91, p_fusion=0.05)
    dump = os.path.join(str(tmp_path), "run.bin")
    write_dump(dump, panel.genes(), b, L)
    m = host.FusionMapper.from_gene_spans(panel.genes(), device=0)
    want = m.scan_pair_end(b)
    m.close()
    raw = b"".join(bytes(r) for r in want)
    out = subprocess.check_output([exe, dump], stderr=subprocess.DEVNULL).decode().strip().split("\n")
    lines = [json.loads(x) for x in out]
    assert len(lines) == 6
    for ln in lines:
        assert ln["records"] == len(want) and ln["pairs"] == b.n, ln
        assert int(ln["checksum"], 16) == fnv(raw), ln
    assert [ln["calls"] for ln in lines] == [70, 2, 1, 70, 2, 1]
