#!/usr/bin/env python3
"""developer tool: end-to-end gf_map_pairs timing (pinned host arenas) under a few settings — where do the milliseconds of
the e2e leg go?  usage (under gpurun): python tools/e2e_probe.py [pairs]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from genefuserust_b200 import synth
from genefuserust_b200._abi import gf_map_stats, gf_match
from genefuserust_b200.host import FusionMapper

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
L = 150
panel = synth.make_panel()
pinned = [torch.empty(P * L, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
batch = synth.generate_pairs(panel, P, read_len=L, seed=12, threads=16, out=tuple(t.numpy() for t in pinned))
off = torch.from_numpy(batch.off1.view(np.int64)).pin_memory()
batch.off1 = off.numpy().view(np.uint64)
batch.off2 = batch.off1
m = FusionMapper.from_gene_spans(panel.genes())
lib, h = m.lib, m.m_indexer.h
cap = max(1 << 16, P // 4)
out = (gf_match * cap)()
n = C.c_uint64(0)


def run(tag, hint, env):
    for k, v in env.items():
        os.environ[k] = v
    hb = batch.as_struct()
    hb.max_len = hint
    for _ in range(2):
        assert lib.gf_map_pairs(h, C.byref(hb), out, cap, C.byref(n)) == 0, lib.gf_last_error()
    ts = []
    for _ in range(12):
        t0 = time.perf_counter()
        lib.gf_map_pairs(h, C.byref(hb), out, cap, C.byref(n))
        ts.append(time.perf_counter() - t0)
    st = gf_map_stats()
    lib.gf_get_map_stats(h, C.byref(st))
    ts.sort()
    print(f"{tag:44s} median {1e3 * ts[6]:7.2f} ms  min {1e3 * ts[0]:7.2f}  device-span {st.ms_total:7.2f} ms  pack {st.ms_host_pack:6.2f} ms (mode {st.packed_upload})  "
          f"h2d {st.h2d_bytes / 1e9:.2f} GB  -> {st.h2d_bytes / ts[6] / 1e9:5.1f} GB/s  {P / ts[6] / 1e6:6.1f} M pairs/s", flush=True)
    for k in env:
        os.environ.pop(k, None)


for rep in range(2):
    run("hybrid upload (default)", 150, {})
    run("every chunk packed", 150, {"GF_HOST_PACK": "1"})
    run("ascii ahead 0 ms (packers only, through the driver thread)", 150, {"GF_ASCII_AHEAD_MS": "0"})
    run("ascii ahead 1 ms", 150, {"GF_ASCII_AHEAD_MS": "1"})
    run("ascii ahead 4 ms", 150, {"GF_ASCII_AHEAD_MS": "4"})
    for t in ("6", "8", "10", "14"):
        run(f"{t} packing threads", 150, {"GF_PACK_THREADS": t, "GF_PACK_MIN_THREADS": "1"})
    for mb in ("128", "384"):
        run(f"chunk {mb} MB", 150, {"GF_CHUNK_MB": mb})
    for parts in ("24", "48"):
        run(f"{parts} parts per packing job", 150, {"GF_PACK_PARTS": parts})
    run("ordinary stores in the packers", 150, {"GF_PACK_NT": "0"})
os.environ["GF_HOST_PACK"] = "0"
run("ASCII upload, hint=150 (per-chunk check), zero-copy qual", 150, {})
run("ASCII upload, hint=0 (pre-scan), zero-copy qual", 0, {})
run("ASCII upload, qualities copied", 150, {"GF_ZEROCOPY_QUAL": "0"})
