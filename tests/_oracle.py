"""ctypes loader for the CPU oracle (oracle/libgf_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from genefuserust_b200._abi import (gf_alignable_result, gf_batch, gf_gene_span, gf_lookup, gf_match, gf_params,
                                    gf_ref_contig)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libgf_oracle.so")


class orc_seqmatch(C.Structure):
    _fields_ = [("seq_start", C.c_int32), ("seq_end", C.c_int32), ("contig", C.c_int32), ("position", C.c_int32)]

    def astuple(self):
        return (self.seq_start, self.seq_end, self.contig, self.position)


_lib = None


def build_oracle():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("gf_oracle.cpp", "gf_oracle_matcher.cpp", "gf_oracle.h")]
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ORACLE_SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build_oracle())
    P = C.POINTER
    L.orc_index_create.argtypes = [P(gf_gene_span), C.c_uint32, P(gf_params)]
    L.orc_index_create.restype = C.c_void_p
    L.orc_index_destroy.argtypes = [C.c_void_p]
    L.orc_index_counts.argtypes = [C.c_void_p, P(C.c_uint64)]
    L.orc_index_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, P(gf_lookup)]
    L.orc_index_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.orc_index_keys.restype = C.c_uint64
    L.orc_map_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, P(orc_seqmatch)]
    L.orc_map_read.restype = C.c_int
    L.orc_fusion_map_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, P(C.c_int), P(gf_match)]
    L.orc_fusion_map_read.restype = C.c_int
    L.orc_fast_merge.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32,
                                 C.c_char_p, C.c_char_p, P(C.c_int32), P(C.c_int32), P(C.c_int32)]
    L.orc_fast_merge.restype = C.c_int
    L.orc_reverse_complement.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
    L.orc_edit_distance.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
    L.orc_edit_distance.restype = C.c_int64
    L.orc_levenshtein_dp.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
    L.orc_levenshtein_dp.restype = C.c_int64
    L.orc_gp_to_i64.argtypes = [C.c_int16, C.c_int32]
    L.orc_gp_to_i64.restype = C.c_int64
    L.orc_i64_to_gp.argtypes = [C.c_int64, P(C.c_int16), P(C.c_int32)]
    L.orc_segment_mask.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_int64, P(orc_seqmatch)]
    L.orc_segment_mask.restype = C.c_int
    L.orc_make_kmer.argtypes = [C.c_char_p, C.c_int32]
    L.orc_make_kmer.restype = C.c_int64
    L.orc_scan_pairs.argtypes = [C.c_void_p, P(gf_batch), P(gf_match), C.c_uint64, C.c_int]
    L.orc_scan_pairs.restype = C.c_uint64
    L.orc_last_scan_counters.argtypes = [P(C.c_uint64)]
    L.orc_get_ref_seq.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
    L.orc_get_ref_seq.restype = C.c_int32
    L.orc_adjust_fusion_break.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32,
                                          P(C.c_int32)]
    L.orc_adjust_fusion_break.restype = C.c_int
    L.orc_bucket_sort.argtypes = [P(gf_match), C.c_uint64, C.c_uint32, P(C.c_char_p), C.c_int, P(C.c_uint64), P(C.c_int64)]
    L.orc_bucket_sort.restype = C.c_uint64
    L.orc_remove_alignables.argtypes = [P(gf_ref_contig), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                        P(gf_alignable_result)]
    L.orc_remove_alignables.restype = C.c_int
    _lib = L
    return L


def make_gene_spans(genes):
    """genes: list of (seq: bytes, reversed: bool).  Returns (array, keepalive)."""
    arr = (gf_gene_span * max(1, len(genes)))()
    keep = []
    for i, (seq, rev) in enumerate(genes):
        buf = C.create_string_buffer(seq, len(seq)) if len(seq) else None
        keep.append(buf)
        arr[i].seq = C.cast(buf, C.c_void_p).value if buf is not None else None
        arr[i].len = len(seq)
        arr[i].reversed = 1 if rev else 0
    return arr, keep


class OracleIndex:
    def __init__(self, genes, params=None):
        self.L = lib()
        self.params = params or gf_params.default()
        arr, keep = make_gene_spans(genes)
        self.h = self.L.orc_index_create(arr, len(genes), C.byref(self.params))
        del keep

    def close(self):
        if self.h:
            self.L.orc_index_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def counts(self):
        out = (C.c_uint64 * 5)()
        self.L.orc_index_counts(self.h, out)
        return dict(zip(("n_sites", "n_keys", "n_unique", "n_normal", "n_high"), map(int, out)))

    def keys(self):
        n = self.L.orc_index_keys(self.h, None, 0)
        a = np.zeros(n, dtype=np.uint32)
        self.L.orc_index_keys(self.h, a.ctypes.data, n)
        return a

    def lookup(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint32)
        out = (gf_lookup * max(1, len(kmers)))()
        self.L.orc_index_lookup(self.h, kmers.ctypes.data, len(kmers), out)
        res = []
        for i in range(len(kmers)):
            o = out[i]
            res.append((o.kind, o.n_sites, tuple((o.contig[j], o.position[j]) for j in range(o.n_sites))))
        return res

    def map_read(self, seq):
        out = (orc_seqmatch * 2)()
        n = self.L.orc_map_read(self.h, seq, len(seq), out)
        return [out[i].astuple() for i in range(n)]

    def fusion_map_read(self, seq):
        m = gf_match()
        mapable = C.c_int(0)
        hit = self.L.orc_fusion_map_read(self.h, seq, len(seq), C.byref(mapable), C.byref(m))
        return (m if hit else None), bool(mapable.value)

    def scan(self, batch, threads=1, cap=None):
        cap = cap or (2 * batch.n + 16)
        out = (gf_match * cap)()
        st = batch.as_struct()
        n = self.L.orc_scan_pairs(self.h, C.byref(st), out, cap, threads)
        assert n <= cap
        return [out[i].astuple() for i in range(n)]

    def counters(self):
        out = (C.c_uint64 * 6)()
        self.L.orc_last_scan_counters(out)
        return dict(zip(("n_mapped", "n_probes1", "n_gated", "n_merged", "seq_bytes", "n_panic"), map(int, out)))


def fast_merge(s1, q1, s2, q2):
    L = lib()
    oseq = C.create_string_buffer(len(s1) + len(s2) + 1)
    oqual = C.create_string_buffer(len(s1) + len(s2) + 1)
    olen, diff, mlen = C.c_int32(), C.c_int32(), C.c_int32()
    ok = L.orc_fast_merge(s1, q1, len(s1), s2, q2, len(s2), oseq, oqual, C.byref(mlen), C.byref(olen), C.byref(diff))
    if not ok:
        return None
    return oseq.raw[:mlen.value], oqual.raw[:mlen.value], olen.value, diff.value


def reverse_complement(s):
    out = C.create_string_buffer(len(s) + 1)
    lib().orc_reverse_complement(s, len(s), out)
    return out.raw[:len(s)]


def edit_distance(a, b):
    return lib().orc_edit_distance(a, len(a), b, len(b))


def levenshtein_dp(a, b):
    return lib().orc_levenshtein_dp(a, len(a), b, len(b))


def segment_mask(mask, gp1, gp2):
    out = (orc_seqmatch * 2)()
    n = lib().orc_segment_mask(bytes(mask), len(mask), gp1, gp2, out)
    return [out[i].astuple() for i in range(n)]


def get_ref_seq(ref, start, end):
    """get_ref_seq (fusion_result.rs:770-798)"""
    out = C.create_string_buffer(len(ref) + 1)
    n = lib().orc_get_ref_seq(ref, len(ref), start, end, out)
    return out.raw[:n]


def adjust_fusion_break(seq, read_break, left_ref, right_ref):
    """FusionResult::adjust_fusion_break for one match -> (shift, left_distance, right_distance, status)"""
    out = (C.c_int32 * 3)()
    st = lib().orc_adjust_fusion_break(seq, len(seq), read_break, left_ref, len(left_ref), right_ref, len(right_ref), out)
    return (0, 0, 0, 1) if st else (out[0], out[1], out[2], 0)


def remove_alignables(contigs, seqs):
    """FusionMapper::remove_alignables (fusion_mapper.rs:488-542) on the CPU: contigs = list of bytes in name order, seqs =
    list of bytes.  Returns (flags, gf_alignable_result, rc) like genefuserust_b200.host.Matcher.remove_alignables."""
    L = lib()
    keep = [np.frombuffer(c, dtype=np.uint8) for c in contigs]
    arr = (gf_ref_contig * max(1, len(contigs)))()
    for i, a in enumerate(keep):
        arr[i].seq = a.ctypes.data if len(a) else None
        arr[i].len = len(a)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=off[1:])
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, np.uint8)
    flags = np.zeros(max(1, len(seqs)), dtype=np.uint8)
    res = gf_alignable_result()
    rc = L.orc_remove_alignables(arr, len(contigs), arena.ctypes.data if len(arena) else None, off.ctypes.data, len(seqs),
                                 flags.ctypes.data, C.byref(res))
    return flags[:len(seqs)], res, rc


def bucket_sort(records, n_genes, names, drop_filtered=True):
    """add_match + per-record filters + sort_matches (fusion_mapper.rs:253-275, 298-385; read_match.rs:203-229).
    records: list of gf_match (push order), names: list of bytes.  Returns [(index into records, bucket), ...]."""
    L = lib()
    n = len(records)
    arr = (gf_match * max(1, n))(*records)
    nm = (C.c_char_p * max(1, n))(*names)
    oi = (C.c_uint64 * max(1, n))()
    ob = (C.c_int64 * max(1, n))()
    k = L.orc_bucket_sort(arr, n, n_genes, nm, 1 if drop_filtered else 0, oi, ob)
    return [(int(oi[i]), int(ob[i])) for i in range(k)]
