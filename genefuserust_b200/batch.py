"""Host-side read batches: byte arenas + (n+1) uint64 offsets, the layout gf_batch describes."""
import ctypes as C

import numpy as np

from ._abi import gf_batch


def _arena(strings):
    lens = np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings))
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(strings), dtype=np.uint8).copy() if len(strings) else np.zeros(0, np.uint8)
    return buf, off


class ReadBatch:
    """A batch of reads (SE) or read pairs (PE).  seq/qual share offsets per mate."""

    def __init__(self, seq1, qual1, off1, seq2=None, qual2=None, off2=None):
        self.seq1 = np.ascontiguousarray(seq1, dtype=np.uint8)
        self.qual1 = np.ascontiguousarray(qual1, dtype=np.uint8)
        self.off1 = np.ascontiguousarray(off1, dtype=np.uint64)
        self.paired = seq2 is not None
        if self.paired:
            self.seq2 = np.ascontiguousarray(seq2, dtype=np.uint8)
            self.qual2 = np.ascontiguousarray(qual2, dtype=np.uint8)
            self.off2 = np.ascontiguousarray(off2, dtype=np.uint64)
            assert len(self.off2) == len(self.off1)
        else:
            self.seq2 = self.qual2 = self.off2 = None
        assert len(self.seq1) == len(self.qual1) == int(self.off1[-1])
        self.n = len(self.off1) - 1
        self.max_len = int(np.diff(self.off1).max()) if self.n else 0
        if self.paired and self.n:
            self.max_len = max(self.max_len, int(np.diff(self.off2).max()))

    @classmethod
    def from_reads(cls, r1, r2=None):
        """r1 / r2: lists of (seq: bytes, qual: bytes)."""
        s1, o1 = _arena([s for s, _ in r1])
        q1, _ = _arena([q for _, q in r1])
        if r2 is None:
            return cls(s1, q1, o1)
        s2, o2 = _arena([s for s, _ in r2])
        q2, _ = _arena([q for _, q in r2])
        return cls(s1, q1, o1, s2, q2, o2)

    def read(self, i, mate=1):
        seq, qual, off = (self.seq1, self.qual1, self.off1) if mate == 1 else (self.seq2, self.qual2, self.off2)
        a, b = int(off[i]), int(off[i + 1])
        return seq[a:b].tobytes(), qual[a:b].tobytes()

    def slice(self, lo, hi):
        def cut(seq, qual, off):
            a, b = int(off[lo]), int(off[hi])
            return seq[a:b], qual[a:b], off[lo:hi + 1] - off[lo]
        s1, q1, o1 = cut(self.seq1, self.qual1, self.off1)
        if not self.paired:
            return ReadBatch(s1, q1, o1)
        s2, q2, o2 = cut(self.seq2, self.qual2, self.off2)
        return ReadBatch(s1, q1, o1, s2, q2, o2)

    def as_struct(self):
        """gf_batch with HOST pointers (the arrays must outlive the call)."""
        b = gf_batch()
        b.n = self.n
        b.seq1 = self.seq1.ctypes.data
        b.qual1 = self.qual1.ctypes.data
        b.off1 = self.off1.ctypes.data
        b.bytes1 = int(self.off1[-1])
        b.max_len = self.max_len
        if self.paired:
            b.seq2 = self.seq2.ctypes.data
            b.qual2 = self.qual2.ctypes.data
            b.off2 = self.off2.ctypes.data
            b.bytes2 = int(self.off2[-1])
        else:
            b.seq2 = b.qual2 = b.off2 = None
            b.bytes2 = 0
        return b

    @property
    def nbytes(self):
        t = self.seq1.nbytes + self.qual1.nbytes + self.off1.nbytes
        if self.paired:
            t += self.seq2.nbytes + self.qual2.nbytes + self.off2.nbytes
        return t
