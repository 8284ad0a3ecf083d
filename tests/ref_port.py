"""A SECOND, independent restatement of the reference's mapping path, in plain Python, written directly from the Rust
sources (not from oracle/gf_oracle.cpp).  Test infrastructure: tests/test_oracle_crosscheck.py runs it against the C++
oracle on randomised cases, so a transcription slip in either restatement shows up as a disagreement.

Follows /root/reference/src/core: indexer.rs:122-250 (make_index / index_contig), :252-538 (map_read), :541-608
(in_required_direction), :616-679 (segment_mask), :689-732; fusion_mapper.rs:93-251; sequence.rs:22-60."""

KMER = 16
CODE = {ord("A"): 0, ord("T"): 1, ord("C"): 2, ord("G"): 3}
COMP = {ord("A"): "T", ord("a"): "T", ord("T"): "A", ord("t"): "A", ord("C"): "G", ord("c"): "G", ord("G"): "C", ord("g"): "C"}


def reverse_complement(s: bytes) -> bytes:
    return "".join(COMP.get(c, "N") for c in reversed(s)).encode()


def make_kmer(seq: bytes, pos: int) -> int:
    k = 0
    for c in seq[pos:pos + KMER]:
        if c not in CODE:
            return -1
        k = (k << 2) | CODE[c]
    return k


def gp_to_i64(contig: int, position: int) -> int:
    return (contig << 32) | (position & 0xFFFFFFFF)


def i64_to_gp(v: int):
    contig = (v >> 32) & 0xFFFF
    if contig >= 0x8000:
        contig -= 0x10000
    pos = v & 0xFFFFFFFF
    if pos >= 0x80000000:
        pos -= 0x100000000
    return contig, pos


class RefIndexer:
    def __init__(self, genes, dup_threshold=5, major=40, minor=20, mismatch=10):
        self.thr, self.major, self.minor, self.mismatch = dup_threshold, major, minor, mismatch
        self.kmer_pos = {}      # kmer -> (contig, position); contig -1 = NORMAL dupe (position = list index), -2 = HIGH
        self.dupe_list = []
        self.fusion_seq = []
        self.reversed = []
        for ctg, (seq, rev) in enumerate(genes):
            self.reversed.append(rev)
            if len(seq) == 0:
                self.fusion_seq.append(b"")
                continue
            s = seq.upper()
            self.index_contig(ctg, s, 0)
            self.index_contig(ctg, reverse_complement(s), 1 - len(s))
            self.fusion_seq.append(s)

    def index_contig(self, ctg, seq, start):
        for i in range(0, len(seq) - KMER):            # exclusive: the window at len-16 is never indexed
            kmer = make_kmer(seq, i)
            if kmer < 0:
                continue
            site = (ctg, i + start)
            gp = self.kmer_pos.get(kmer)
            if gp is None:
                self.kmer_pos[kmer] = site
            elif gp[0] == -2:
                continue
            elif gp[0] == -1:
                if len(self.dupe_list[gp[1]]) >= self.thr:
                    self.kmer_pos[kmer] = (-2, gp[1])
                    self.dupe_list[gp[1]] = []
                else:
                    self.dupe_list[gp[1]].append(site)
            else:
                self.dupe_list.append([gp, site])
                self.kmer_pos[kmer] = (-1, len(self.dupe_list) - 1)

    def sites(self, kmer):
        gp = self.kmer_pos.get(kmer)
        if gp is None or gp[0] == -2:
            return []
        if gp[0] == -1:
            return list(self.dupe_list[gp[1]])
        return [gp]

    def map_read(self, seq: bytes):
        seqlen = len(seq)
        stat = {0: 0}
        for i in range(0, seqlen - KMER + 1, 2):
            kmer = make_kmer(seq, i)
            if kmer < 0:
                continue
            if kmer not in self.kmer_pos:
                stat[0] += 1
                continue
            for (c, p) in self.sites(kmer):
                g = gp_to_i64(c, p - i)
                stat[g] = stat.get(g, 0) + 1
        gp1 = gp2 = 0
        count1 = count2 = 0
        for k in sorted(stat):                          # BTreeMap iteration order
            v = stat[k]
            if k != 0 and v > count1:
                gp2, count2 = gp1, count1
                gp1, count1 = k, v
            elif k != 0 and v > count2:
                gp2, count2 = k, v
        if count1 * 2 < self.major or count2 * 2 < self.minor:
            return []
        mask = [0] * seqlen
        for i in range(0, seqlen - KMER + 1):
            kmer = make_kmer(seq, i)
            if kmer < 0 or kmer not in self.kmer_pos:
                continue
            for (c, p) in self.sites(kmer):
                g = gp_to_i64(c, p - i)
                if abs(g - gp1) <= 1:
                    flag = 3
                elif abs(g - gp2) <= 1:
                    flag = 2
                elif g == 0:
                    flag = 1
                else:
                    continue
                for q in range(i, min(seqlen, i + KMER)):
                    mask[q] = max(mask[q], flag)
        if sum(1 for m in mask if m in (0, 1)) > self.mismatch:
            return []
        return segment_mask(mask, seqlen, i64_to_gp(gp1), i64_to_gp(gp2))

    def in_required_direction(self, mapping):
        if len(mapping) < 2:
            return False
        left, right = mapping[0], mapping[1]
        if left[0] > right[0]:
            left, right = right, left
        lp, rp = left[2][1], right[2][1]
        if lp > 0 and rp > 0:
            return True
        if lp < 0 and rp < 0:
            return False
        lrev, rrev = self.reversed[left[2][0]], self.reversed[right[2][0]]
        if lrev and not rrev:
            return False
        if not lrev and rrev:
            return True
        if left[2][0] < right[2][0]:
            return True
        return False                                    # "left < left" typo: never true

    def calc_ed(self, seq, contig, start, end):
        if (start >= 0 and end <= 0) or (start <= 0 and end >= 0):
            return -1
        fs = self.fusion_seq[contig]
        if abs(start) >= len(fs) or abs(end) >= len(fs):
            return -2
        ss = seq
        if start < 0:
            ss = reverse_complement(seq)
            start, end = -end, -start
        return levenshtein(ss, fs[start:end + 1])

    def fusion_map_read(self, seq):
        """-> (match tuple or None, mapable)"""
        mapping = self.map_read(seq)
        if len(mapping) < 2:
            return None, False
        if not self.in_required_direction(mapping):
            return None, True
        left, right = mapping[0], mapping[1]
        if left[0] > right[0]:
            left, right = right, left
        read_break = (left[1] + right[0]) // 2
        lc, lp = left[2][0], left[2][1] + read_break
        rc, rp = right[2][0], right[2][1] + read_break + 1
        gap = right[0] - left[1] - 1
        left_len, right_len = read_break + 1, len(seq) - (read_break + 1)
        ld = self.calc_ed(seq[:left_len], lc, lp - left_len + 1, lp)
        rd = self.calc_ed(seq[read_break + 1:], rc, rp, rp + right_len - 1)
        return (read_break, lc, lp, rc, rp, gap, ld, rd, len(seq)), True


def segment_mask(mask, seqlen, gp1, gp2):
    result = []
    for target, gp in ((3, gp1), (2, gp2)):
        max_start = max_end = -1
        start = 0
        while True:
            while mask[start] != target and start != seqlen - 1:
                start += 1
            if start >= seqlen - 1:
                break
            if mask[start] == target:
                end = start + 1
                g = 0
                while g < 10 and end + g < seqlen:
                    if mask[end + g] > target:
                        break
                    if mask[end + g] == target:
                        end += g + 1
                        g = 0
                        continue
                    g += 1
                end -= 1
                if end - start > max_end - max_start:
                    max_end, max_start = end, start
                start += 1
            else:
                break
        if max_end - max_start > 20:
            result.append((max_start, max_end, gp))
    return result


def levenshtein(a: bytes, b: bytes) -> int:
    if not a:
        return len(b)
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def fast_merge(s1: bytes, q1: bytes, s2: bytes, q2: bytes):
    """SequenceReadPair::fast_merge (read.rs:313-440) -> (merged_seq, olen, diff) or None"""
    str2, qual2 = reverse_complement(s2), q2[::-1]
    len1, len2 = len(s1), len(s2)
    olen, overlapped, diff = 30, False, 0
    while olen <= min(len1, len2):
        diff = low = 0
        ok = True
        offset = len1 - olen
        for i in range(olen):
            if s1[offset + i] != str2[i]:
                diff += 1
                a, b = q1[offset + i], qual2[i]
                if (a >= ord("?") and b <= ord("0")) or (a <= ord("0") and b >= ord("?")):
                    low += 1
                if diff > low or low >= 3:
                    ok = False
                    break
        if ok:
            overlapped = True
            break
        olen += 1
    if not overlapped:
        return None
    offset = len1 - olen
    merged = bytearray(s1[:offset] + str2)
    for i in range(olen):
        if s1[offset + i] != str2[i]:
            if q1[offset + i] >= ord("?") and qual2[i] <= ord("0"):
                merged[offset + i] = s1[offset + i]
            else:
                merged[offset + i] = str2[i]
    return bytes(merged), olen, diff


def scan_pair(ix: "RefIndexer", s1, q1, s2, q2):
    """PairEndScanner::scan_pair_end for one pair (pescanner.rs:427-518) -> list of
    (source, used_rc, reversed, read_break, lc, lp, rc, rp, gap, ld, rd, seq_len)"""
    out = []

    def try_read(seq, source, set_reversed):
        m, mapable = ix.fusion_map_read(seq)
        if m is not None:
            out.append((source, 0, 0) + m)
        elif mapable:
            m2, _ = ix.fusion_map_read(reverse_complement(seq))
            if m2 is not None:
                out.append((source, 1, 1 if set_reversed else 0) + m2)
    merged = fast_merge(s1, q1, s2, q2)
    if merged is not None:
        try_read(merged[0], 0, False)       # merged rc matches are NOT flagged reversed (:455-469)
        return out
    try_read(s1, 1, True)
    try_read(s2, 2, True)
    return out


# ---- report stage: src/core/fusion_result.rs (second, independent restatement; the C++ one is oracle/gf_oracle.cpp) ----
def get_ref_seq(ref: bytes, start: int, end: int) -> bytes:
    """fusion_result.rs:770-798"""
    if (start >= 0 and end <= 0) or (start <= 0 and end >= 0):
        return b""
    if abs(start) >= len(ref) or abs(end) >= len(ref):
        return b""
    n = abs(end - start) + 1
    if start < 0:
        return reverse_complement(ref[-end:-end + n])
    return ref[start:start + n]


def adjust_fusion_break(seq: bytes, read_break: int, left_ref: bytes, right_ref: bytes):
    """FusionResult::adjust_fusion_break + calc_ed (fusion_result.rs:299-397) for one match ->
    (shift, left_distance, right_distance), or None where a shifted break leaves the read"""
    best, shift, out_l, out_r = 0xFFFF, 0, 0, 0
    for s in range(-3, 4):
        left_len = read_break + s + 1
        right_len = len(seq) - left_len
        if left_len < 0 or right_len < 0:
            return None
        left_seq, right_seq = seq[:left_len], seq[left_len:]
        lc = min(len(left_seq), len(left_ref), 20)
        rc = min(len(right_seq), len(right_ref), 20)
        total = levenshtein(left_seq[len(left_seq) - lc:], left_ref[len(left_ref) - lc:]) + \
            levenshtein(right_seq[:rc], right_ref[:rc])
        lc = min(left_len, len(left_ref))
        rc = min(right_len, len(right_ref))
        le = levenshtein(left_seq[len(left_seq) - lc:], left_ref[len(left_ref) - lc:])
        re_ = levenshtein(right_seq[:rc], right_ref[:rc])
        if total < best:
            best, shift, out_l, out_r = total, s, le, re_
    return shift, out_l, out_r
