"""Synthetic workloads of SURVEY.md §8(d): a cancer.csv-shaped gene panel on random contigs with planted
repeats and fusions, and counter-based read pairs.  Host utility for tests and bench (not the product path)."""
import ctypes as C
import os
import subprocess

import numpy as np

from ..batch import ReadBatch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgf_synth.so")
SRC = os.path.join(HERE, "gf_synth.cpp")
GENE_TABLE = os.path.join(os.path.dirname(HERE), "data", "cancer_genes.tsv")


class gfs_fusion(C.Structure):
    _fields_ = [("gene_a", C.c_int32), ("pos_a", C.c_int32), ("strand_a", C.c_int32),
                ("gene_b", C.c_int32), ("pos_b", C.c_int32), ("strand_b", C.c_int32)]


class gfs_config(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_int32), ("p_target", C.c_double), ("p_fusion", C.c_double),
                ("sub_rate", C.c_double), ("n_rate", C.c_double), ("n_genes", C.c_int32),
                ("gene_arena", C.c_void_p), ("gene_off", C.c_void_p), ("n_fusions", C.c_int32),
                ("fusions", C.c_void_p)]


_lib = None


def build():
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-pthread", "-shared", "-o", SO, SRC])
    return SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.gfs_random_bases.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64]
        L.gfs_generate_pairs.argtypes = [C.POINTER(gfs_config), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def random_bases(seed, n):
    out = np.empty(n, dtype=np.uint8)
    lib().gfs_random_bases(seed, out.ctypes.data, n)
    return out


def load_gene_table(path=GENE_TABLE):
    """[(name, chr, start, end, reversed, n_exons)] — the shape of testdata/cancer.csv."""
    rows = []
    for line in open(path):
        if line.startswith("#"):
            continue
        name, chrom, a, b, rev, nex = line.rstrip("\n").split("\t")
        rows.append((name, chrom, int(a), int(b), int(rev), int(nex)))
    return rows


class Panel:
    """Gene panel: per-gene upper-case sequence + reversed flag + planted fusions."""

    def __init__(self, names, seqs, reversed_flags, fusions):
        self.names = names
        self.seqs = seqs                  # list of np.uint8 arrays
        self.reversed = reversed_flags    # list of 0/1
        self.fusions = fusions            # list of gfs_fusion-like tuples
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        self.off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=self.off[1:])
        self.arena = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)

    @property
    def n_genes(self):
        return len(self.seqs)

    def genes(self):
        """[(seq bytes, reversed)] in CSV order — what Indexer::make_index slices out."""
        return [(s.tobytes(), bool(r)) for s, r in zip(self.seqs, self.reversed)]


def make_panel(seed=20240201, scale=1.0, max_genes=None, n_fusions=20, plant=True, repeat_frac=0.0):
    """cancer.csv-shaped panel.  `scale` shrinks every gene length (tests); 1.0 = 15.1 Mbases.  repeat_frac > 0 plants
    blocks of 300..3000 bases in 2..5 copies each (either strand) over random genes until that fraction of the panel's
    bases belongs to a copy: NORMAL dupe keys, the shape of interspersed repeats in real genes."""
    table = load_gene_table()
    if max_genes:
        table = table[:max_genes]
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    seqs, names, revs = [], [], []
    for gi, (name, _chr, a, b, rev, _nex) in enumerate(table):
        n = max(200, int((b - a) * scale))
        seqs.append(random_bases(seed * 1000003 + gi, n))
        names.append(name)
        revs.append(rev)
    if plant and len(seqs) >= 8:
        def put(gi, pos, block):
            s = seqs[gi]
            pos = min(pos, max(0, len(s) - len(block)))
            s[pos:pos + len(block)] = block[:len(s) - pos]
        blk = random_bases(seed + 77, 2000)
        for gi in (0, 3, 5):                       # NORMAL dupes: 3 copies (fwd) -> 3 sites per k-mer
            put(gi, len(seqs[gi]) // 3, blk)
        blk = random_bases(seed + 78, 300)
        for k in range(8):                         # HIGH dupes: 8 copies
            put(k % len(seqs), 50 + 400 * k, blk)
        put(1, len(seqs[1]) // 2, np.frombuffer(b"AT" * 200, dtype=np.uint8))
        put(2, len(seqs[2]) // 2, np.frombuffer(b"A" * 100, dtype=np.uint8))
        put(4, len(seqs[4]) // 2, np.frombuffer(b"N" * 500, dtype=np.uint8))
    if repeat_frac > 0:
        rr = np.random.RandomState((seed + 4242) & 0x7FFFFFFF)
        total = sum(len(s) for s in seqs)
        planted = 0
        comp = np.zeros(256, dtype=np.uint8)
        for a_, b_ in zip(b"ACGTN", b"TGCAN"):
            comp[a_] = b_
        k = 0
        while planted < repeat_frac * total:
            blk = random_bases(seed + 100000 + k, int(rr.randint(300, 3001)))
            k += 1
            for _ in range(int(rr.randint(2, 6))):
                gi = int(rr.randint(0, len(seqs)))
                s = seqs[gi]
                if len(s) <= len(blk) + 2:
                    continue
                pos = int(rr.randint(0, len(s) - len(blk)))
                s[pos:pos + len(blk)] = blk if rr.rand() < 0.5 else comp[blk[::-1]]
                planted += len(blk)
    fusions = []
    ng = len(seqs)
    for k in range(n_fusions):
        ga = int(rng.randint(0, ng))
        gb = int(rng.randint(0, ng))
        while gb == ga and ng > 1:
            gb = int(rng.randint(0, ng))
        la, lb = len(seqs[ga]), len(seqs[gb])
        pa = int(rng.randint(la // 4, 3 * la // 4))
        pb = int(rng.randint(lb // 4, 3 * lb // 4))
        sa = 1 if (k & 1) == 0 else -1
        sb = 1 if (k & 2) == 0 else -1
        fusions.append((ga, pa, sa, gb, pb, sb))
    return Panel(names, seqs, revs, fusions)


def generate_pairs(panel, n, read_len=150, seed=12, first=0, p_target=0.70, p_fusion=0.001, sub_rate=0.002,
                   n_rate=0.0005, threads=None, out=None, return_kind=False):
    """Counter-based read pairs [first, first+n).  `out` = optional preallocated (seq1, qual1, seq2, qual2)
    uint8 arrays of n*read_len bytes (e.g. views of pinned torch tensors)."""
    threads = threads or min(32, os.cpu_count() or 1)
    L = read_len
    if out is None:
        out = tuple(np.empty(n * L, dtype=np.uint8) for _ in range(4))
    s1, q1, s2, q2 = out
    fus = (gfs_fusion * max(1, len(panel.fusions)))()
    for i, f in enumerate(panel.fusions):
        fus[i] = gfs_fusion(*f)
    cfg = gfs_config(seed, L, p_target, p_fusion, sub_rate, n_rate, panel.n_genes, panel.arena.ctypes.data,
                     panel.off.ctypes.data, len(panel.fusions), C.cast(fus, C.c_void_p).value)
    kind = np.empty(n, dtype=np.uint8) if return_kind else None
    lib().gfs_generate_pairs(C.byref(cfg), first, n, s1.ctypes.data, q1.ctypes.data, s2.ctypes.data,
                             q2.ctypes.data, kind.ctypes.data if return_kind else None, threads)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    batch = ReadBatch(s1, q1, off, s2, q2, off.copy())
    return (batch, kind) if return_kind else batch
