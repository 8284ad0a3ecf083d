"""Test-only report driver: a Python restatement of what the reference does with the
match records AFTER the hot path, so that report-level parity — "same unique read counts per fusion, byte-identical fusion
reports" (BASELINE.json north_star) — can be demonstrated with device results feeding the report stage:

  FusionMapper::filter_matches / sort_matches / cluster_matches      src/core/fusion_mapper.rs:276-557
  FusionResult (support, calc_fusion_point, make_reference, adjust_fusion_break, calc_unique, update_info, is_qualified,
                is_deletion, protein direction)                       src/core/fusion_result.rs:50-410, 770-798
  Gene::pos2str / get_exon_intron / gene_pos_2_chr_pos, Gene::parse   src/core/gene.rs:44-215
  Fusion::parse_csv                                                   src/core/fusion.rs:23-91
  JsonReporter::run                                                   src/core/json_reporter.rs:34-112

TEST INFRASTRUCTURE, NOT PRODUCT, and "parity unpinned" like the oracle: the reference cannot be built here and holds no
test for these stages.  In the product the host Rust keeps all of this; the driver exists so that the two places where the
library can take work from that stage (device filter / order keys, gf_adjust_fusion_break, gf_alignable_filter) are checked
end to end against the all-CPU pipeline.  Places where the Rust code would panic raise RefPanic.
"""
from dataclasses import dataclass, field

FUSIONSCAN_VER = "0.1.2"            # src/core/html_reporter.rs (FUSIONSCAN_VER), the value in the JSON header


class RefPanic(Exception):
    pass


# ------------------------------------------------------------------------------------------------ gene.rs / fusion.rs
@dataclass
class Exon:
    id: int
    start: int
    end: int


@dataclass
class Gene:
    m_name: str = "invalid"
    m_chr: str = "invalid"
    m_start: int = 0
    m_end: int = 0
    m_exons: list = field(default_factory=list)
    m_reversed: bool = False

    def is_reversed(self):
        return self.m_reversed

    def valid(self):                                   # gene.rs:40-42
        return self.m_name != "invalid" and self.m_start != 0 and self.m_end != 0

    @staticmethod
    def parse(line_str):                               # gene.rs:44-90
        sp = line_str.split(",")
        if len(sp) < 2:
            return Gene()
        name = sp[0][1:].strip()
        chr_pos = sp[1].split(":")
        if len(chr_pos) < 2:
            return Gene()
        rng = chr_pos[1].split("-")
        if len(rng) < 2:
            return Gene()
        return Gene(name, chr_pos[0].strip(), int(rng[0].strip()), int(rng[1].strip()))

    def add_exon(self, id_, start, end):               # gene.rs:92-107
        self.m_exons.append(Exon(id_, start, end))
        if len(self.m_exons) > 1 and self.m_exons[0].start > self.m_exons[1].start:
            self.m_reversed = True

    def pos2str(self, pos):                            # gene.rs:134-175
        pp = abs(pos) + self.m_start
        ss = f"{self.m_name}:"
        ex = self.m_exons
        for i in range(len(ex)):
            if ex[i].start <= pp <= ex[i].end:
                ss += f"exon:{ex[i].id}|"
                break
            if i > 0:
                if self.m_reversed:
                    if ex[i].end < pp < ex[i - 1].start:
                        ss += f"intron:{ex[i].id - 1}|"
                        break
                else:
                    if ex[i - 1].end < pp < ex[i].start:
                        ss += f"intron:{ex[i].id - 1}|"
                        break
        ss += "+" if pos >= 0 else "-"
        ss += f"{self.m_chr}:{pp}"
        return ss

    def get_exon_intron(self, pos, is_exon, number):   # gene.rs:177-206; returns the updated (is_exon, number)
        pp = abs(pos) + self.m_start
        prev_exon = self.m_exons[0] if self.m_exons else None   # `prev_exon` is never advanced in the reference
        for i, exon in enumerate(self.m_exons):
            if exon.start <= pp <= exon.end:
                return True, exon.id
            if i > 0:
                if self.m_reversed:
                    if exon.end < pp < prev_exon.start:
                        return False, exon.id - 1
                else:
                    if prev_exon.end < pp < exon.start:
                        return False, exon.id - 1
        return is_exon, number

    def gene_pos_2_chr_pos(self, genepos):             # gene.rs:208-215
        chrpos = abs(genepos) + self.m_start
        return -chrpos if genepos < 0 else chrpos


def parse_csv(path):
    """Fusion::parse_csv (fusion.rs:23-91) -> [Gene] (a Fusion is just its gene)"""
    fusions = []
    working = Gene()
    with open(path, "r") as f:
        for raw in f:
            line_str = raw.strip()
            sp = line_str.split(",")
            if len(sp) < 2:
                continue
            if sp[0].startswith("#"):
                continue
            if sp[0].startswith(">"):
                if working.valid():
                    fusions.append(working)
                working = Gene.parse(line_str)
                continue
            if len(sp) < 3:
                continue
            working.add_exon(int(sp[0].strip()), int(sp[1].strip()), int(sp[2].strip()))
    if working.valid():
        fusions.append(working)
    return fusions


# ------------------------------------------------------------------------------------------------ reads / matches
_COMP = {ord("A"): "T", ord("a"): "T", ord("T"): "A", ord("t"): "A", ord("C"): "G", ord("c"): "G", ord("G"): "C", ord("g"): "C"}


def reverse_complement(seq: bytes) -> bytes:           # sequence.rs:22-60
    return "".join(_COMP.get(b, "N") for b in reversed(seq)).encode()


@dataclass
class Read:                                            # read.rs SequenceRead
    m_name: bytes
    m_seq: bytes
    m_strand: bytes
    m_quality: bytes

    def reverse_complement(self):                      # read.rs:243-261
        return Read(self.m_name, reverse_complement(self.m_seq), b"-" if self.m_strand == b"+" else b"+", self.m_quality[::-1])

    def __len__(self):
        return len(self.m_seq)


@dataclass
class GenePos:
    contig: int
    position: int


@dataclass
class ReadMatch:                                       # read_match.rs:17-54
    m_read: Read
    m_read_break: int
    m_left_gp: GenePos
    m_right_gp: GenePos
    m_gap: int
    m_reversed: bool
    m_left_distance: int = 0
    m_right_distance: int = 0
    filter_flags: int = 0        # what the three per-record filters decide (device or oracle)
    push_key: tuple = ()         # (pair_idx, source): the push order at -t 1


def read_of_record(rec, batch, names, fast_merge):
    """the SequenceRead a record's ReadMatch holds: merged read (named '{R1 name} merged_diff_{N}', strand '+',
    read.rs:369-437), R1 or R2, reverse-complemented when the match came from the rc retry (pescanner.rs:455-514).
    `fast_merge(s1, q1, s2, q2)` -> (seq, qual, olen, diff) or None — the host's own fast_merge."""
    i = rec.pair_idx
    s1, q1 = batch.read(i, 1)
    if rec.source == 1:
        rd = Read(names[0][i], s1, b"+", q1)
    else:
        s2, q2 = batch.read(i, 2)
        if rec.source == 2:
            rd = Read(names[1][i], s2, b"+", q2)
        else:
            mg = fast_merge(s1, q1, s2, q2)
            assert mg is not None and (mg[2], mg[3]) == (rec.merge_olen, rec.merge_diff), "record and host fast_merge disagree"
            rd = Read(names[0][i] + b" merged_diff_%d" % mg[3], mg[0], b"+", mg[1])
    if rec.used_rc:
        rd = rd.reverse_complement()
    assert len(rd) == rec.seq_len
    return rd


def make_read_match(rec, rd):
    return ReadMatch(rd, rec.read_break, GenePos(rec.l_contig, rec.l_pos), GenePos(rec.r_contig, rec.r_pos), rec.gap,
                     bool(rec.reversed), rec.l_dist, rec.r_dist, rec.filter_flags, (rec.pair_idx, rec.source))


# ------------------------------------------------------------------------------------------------ filters (fusion_mapper.rs:298-377)
def dis_connected_count(s: bytes) -> int:              # src/utils/mod.rs:48-56 (s.len() - 1 on usize: empty string wraps)
    if len(s) == 0:
        raise RefPanic("dis_connected_count on an empty string")
    return sum(1 for i in range(len(s) - 1) if s[i] != s[i + 1])


def is_low_complexity(s: bytes) -> bool:               # fusion_mapper.rs:559-569
    return len(s) < 20 or dis_connected_count(s) < 7


def per_record_filter_flags(rm: ReadMatch, deletion_threshold=50) -> int:
    """remove_by_complexity | remove_by_distance | remove_indels as the bits of gf_match.filter_flags"""
    seq, rb = rm.m_read.m_seq, rm.m_read_break
    f = 0
    if is_low_complexity(seq[:rb + 1]) or is_low_complexity(seq[rb + 1:]):
        f |= 1
    if rm.m_left_distance + rm.m_right_distance >= 5:
        f |= 2
    if rm.m_left_gp.contig == rm.m_right_gp.contig and abs(rm.m_left_gp.position - rm.m_right_gp.position) < deletion_threshold:
        f |= 4
    return f


def read_match_cmp(a: ReadMatch, b: ReadMatch) -> int:
    """a.partial_cmp(b) (read_match.rs:203-229)"""
    if a.m_read_break != b.m_read_break:
        return -1 if a.m_read_break < b.m_read_break else 1
    la, lb = len(a.m_read.m_seq), len(b.m_read.m_seq)
    if lb != la:                                       # other.len().partial_cmp(&self.len())
        return -1 if lb < la else 1
    if a.m_read.m_name != b.m_read.m_name:
        return -1 if a.m_read.m_name < b.m_read.m_name else 1
    return 0


def sort_matches(fusion_matches):                      # fusion_mapper.rs:379-385
    import functools
    for v in fusion_matches.values():
        v.sort(key=functools.cmp_to_key(lambda a, b: read_match_cmp(b, a)))


# ------------------------------------------------------------------------------------------------ fusion_result.rs
def get_ref_seq(ref_s: bytes, start: int, end: int) -> bytes:      # fusion_result.rs:770-798
    if (start >= 0 and end <= 0) or (start <= 0 and end >= 0):
        return b""
    if abs(start) >= len(ref_s) or abs(end) >= len(ref_s):
        return b""
    ln = abs(end - start) + 1
    if start < 0:
        return reverse_complement(ref_s[-end:-end + ln])
    return ref_s[start:start + ln]


@dataclass
class FusionResult:
    m_left_gp: GenePos = field(default_factory=lambda: GenePos(0, 0))
    m_right_gp: GenePos = field(default_factory=lambda: GenePos(0, 0))
    m_matches: list = field(default_factory=list)
    m_unique: int = 0
    m_title: str = ""
    m_left_ref: bytes = b""
    m_right_ref: bytes = b""
    m_left_ref_ext: bytes = b""
    m_right_ref_ext: bytes = b""
    m_left_pos: str = ""
    m_right_pos: str = ""
    m_left_gene: Gene = None
    m_right_gene: Gene = None
    m_left_is_exon: bool = False
    m_right_is_exon: bool = False
    m_left_exon_or_intron_id: int = -1
    m_right_exon_or_intron_id: int = -1

    def support(self, m):                              # :414-445
        for rm in self.m_matches:
            if (abs(m.m_left_gp.position - rm.m_left_gp.position) <= 3 and abs(m.m_right_gp.position - rm.m_right_gp.position) <= 3
                    and m.m_left_gp.contig == rm.m_left_gp.contig and m.m_right_gp.contig == rm.m_right_gp.contig):
                return True
        return False

    def calc_fusion_point(self):                       # :61-88
        if not self.m_matches:
            return
        lt = rt = 0
        for rm in self.m_matches:
            if rm.m_gap == 0:
                self.m_left_gp = GenePos(rm.m_left_gp.contig, rm.m_left_gp.position)
                self.m_right_gp = GenePos(rm.m_right_gp.contig, rm.m_right_gp.position)
                return
            lt += rm.m_left_gp.position
            rt += rm.m_right_gp.position
        n = len(self.m_matches)
        trunc = lambda x: int(x / n) if x >= 0 else -int(-x / n)       # i64 division truncates toward zero
        self.m_left_gp = GenePos(self.m_matches[0].m_left_gp.contig, _tdiv(lt, n))
        self.m_right_gp = GenePos(self.m_matches[0].m_right_gp.contig, _tdiv(rt, n))

    def make_reference(self, ref_l, ref_r):            # :242-297
        ll = lr = 0
        for rm in self.m_matches:
            ll = max(ll, rm.m_read_break + 1)
            lr = max(lr, len(rm.m_read) - (rm.m_read_break + 1))
        lp, rp = self.m_left_gp.position, self.m_right_gp.position
        self.m_left_ref = get_ref_seq(ref_l, lp - ll + 1, lp)
        self.m_right_ref = get_ref_seq(ref_r, rp, rp + lr - 1)
        self.m_left_ref_ext = get_ref_seq(ref_l, lp, lp + lr - 1)
        self.m_right_ref_ext = get_ref_seq(ref_r, rp - ll + 1, rp)

    def calc_unique(self):                             # :90-108
        self.m_unique = 1
        for prev, mm in zip(self.m_matches, self.m_matches[1:]):
            if mm.m_read_break != prev.m_read_break or len(mm.m_read) != len(prev.m_read):
                self.m_unique += 1

    def is_deletion(self):                             # :110-121
        if self.m_left_gp.contig == self.m_right_gp.contig:
            if self.m_left_gp.position > 0 and self.m_right_gp.position > 0:
                return True
            if self.m_left_gp.position < 0 and self.m_right_gp.position < 0:
                return True
        return False

    def can_be_matched(self, s1, s2, edit_distance):   # :134-166
        ln = len(s1)
        for offset in range(-6, 7):
            start1, start2 = max(offset, 0), max(-offset, 0)
            cmplen = ln - abs(offset)
            if start1 >= len(s1) or start2 >= len(s2):
                return True
            if cmplen < 0 or start1 + cmplen > len(s1) or start2 + cmplen > len(s2):
                raise RefPanic("subchars out of range in can_be_matched")
            ed = edit_distance(s1[start1:start1 + cmplen], s2[start2:start2 + cmplen])
            if ed <= _tdiv(cmplen, 10):
                return True
        return False

    def can_be_mapped(self, edit_distance):            # :123-132
        return (self.can_be_matched(self.m_left_ref_ext, self.m_right_ref, edit_distance)
                or self.can_be_matched(self.m_left_ref, self.m_right_ref_ext, edit_distance))

    def is_qualified(self, edit_distance, unique_requirement=2):     # :168-205
        if self.m_unique < unique_requirement:
            return False
        if self.can_be_mapped(edit_distance):
            return False
        if len(self.m_left_ref) <= 30 or len(self.m_right_ref) <= 30:
            return False
        if dis_connected_count(self.m_left_ref[len(self.m_left_ref) - 10:]) <= 2:
            return False
        if dis_connected_count(self.m_right_ref[:10]) <= 2:
            return False
        return True

    def update_info(self, fusions):                    # :207-240
        self.m_left_gene = fusions[self.m_left_gp.contig]
        self.m_right_gene = fusions[self.m_right_gp.contig]
        ss = "Deletion: " if self.is_deletion() else "Fusion: "
        ss += (f"{self.m_left_gene.pos2str(self.m_left_gp.position)}___{self.m_right_gene.pos2str(self.m_right_gp.position)}"
               f"  (total: {len(self.m_matches)}, unique:{self.m_unique})")
        self.m_title = ss
        self.m_left_pos = self.m_left_gene.pos2str(self.m_left_gp.position)
        self.m_right_pos = self.m_right_gene.pos2str(self.m_right_gp.position)
        self.m_left_is_exon, self.m_left_exon_or_intron_id = self.m_left_gene.get_exon_intron(
            self.m_left_gp.position, self.m_left_is_exon, self.m_left_exon_or_intron_id)
        self.m_right_is_exon, self.m_right_exon_or_intron_id = self.m_right_gene.get_exon_intron(
            self.m_right_gp.position, self.m_right_is_exon, self.m_right_exon_or_intron_id)

    def is_left_protein_forward(self):                 # :446-452
        return self.m_left_gp.position < 0 if self.m_left_gene.is_reversed() else self.m_left_gp.position > 0

    def is_right_protein_forward(self):                # :454-460
        return self.m_right_gp.position < 0 if self.m_right_gene.is_reversed() else self.m_right_gp.position > 0


def _tdiv(a, b):
    """integer division truncating toward zero (Rust `/` on signed integers)"""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def cluster_matches(fusion_matches, n_buckets, fusions, m_fusion_seq, edit_distance, adjust_batch,
                    output_deletions=False, output_untranslated=False):
    """FusionMapper::cluster_matches (fusion_mapper.rs:394-486) + sort_fusion_results (:544-557).
    fusion_matches: {bucket index: [ReadMatch]} already filtered and sorted.
    adjust_batch([(left_ref, right_ref, [(seq, read_break), ...]), ...]) -> [[(shift, left_distance, right_distance), ...], ...]
    is FusionResult::adjust_fusion_break's arithmetic (fusion_result.rs:299-397) for ALL results of the run at once: the CPU
    pipeline passes the oracle's, the device pipeline gf_adjust_fusion_break's."""
    pending = []
    for i in range(n_buckets):
        fm = fusion_matches.get(i, [])
        frs = []
        for rm in fm:
            for fr in frs:
                if fr.support(rm):
                    fr.m_matches.append(rm)
                    break
            else:
                fr = FusionResult()
                fr.m_matches.append(rm)
                frs.append(fr)
        for fr in frs:
            fr.calc_fusion_point()
            fr.make_reference(m_fusion_seq[fr.m_left_gp.contig], m_fusion_seq[fr.m_right_gp.contig])
            pending.append(fr)
    shifts = adjust_batch([(fr.m_left_ref, fr.m_right_ref, [(rm.m_read.m_seq, rm.m_read_break) for rm in fr.m_matches])
                           for fr in pending])
    results = []
    for fr, sh in zip(pending, shifts):
        for rm, (shift, ld, rd) in zip(fr.m_matches, sh):       # fusion_result.rs:303-320
            rm.m_left_distance, rm.m_right_distance = ld, rd
            rm.m_read_break += shift
            rm.m_left_gp.position += shift
            rm.m_right_gp.position += shift
        fr.calc_unique()
        fr.update_info(fusions)
        if fr.is_qualified(edit_distance):
            if not output_deletions and fr.is_deletion():
                continue
            if fr.is_left_protein_forward() != fr.is_right_protein_forward() and not output_untranslated:
                continue
            results.append(fr)
    # sort_by(|a, b| more_reads(b, a)): descending by (m_unique, matches), stable
    import functools

    def more_reads(r1, r2):
        if r1.m_unique != r2.m_unique:
            return -1 if r1.m_unique < r2.m_unique else 1
        if len(r1.m_matches) != len(r2.m_matches):
            return -1 if len(r1.m_matches) < len(r2.m_matches) else 1
        return 0
    results.sort(key=functools.cmp_to_key(lambda a, b: more_reads(b, a)))
    return results


def json_report(results, command="genefuse", time_str="1970-01-01 00:00:00", output_deletions=False, output_untranslated=False):
    """JsonReporter::run (json_reporter.rs:34-112) -> bytes.  `command` and `time_str` are inputs of the writer (COMMAND,
    Local::now())."""
    o = []
    w = o.append
    w("{\n")
    w(f"\t\"command\":\"{command}\",\n")
    w(f"\t\"version\":\"{FUSIONSCAN_VER}\",\n")
    w(f"\t\"time\":\"{time_str}\",\n")
    w("\t\"fusions\":{")
    first = True
    for fusion in results:
        if not output_deletions and fusion.is_deletion():
            continue
        if fusion.is_left_protein_forward() != fusion.is_right_protein_forward() and not output_untranslated:
            continue
        if first:
            w("\n")
            first = False
        else:
            w(",\n")
        w(f"\t\t\"{fusion.m_title}\":{{\n")
        for side, gene, gp, ref, ext, pos, is_exon, eid, fwd in (
                ("left", fusion.m_left_gene, fusion.m_left_gp, fusion.m_left_ref, fusion.m_left_ref_ext, fusion.m_left_pos,
                 fusion.m_left_is_exon, fusion.m_left_exon_or_intron_id, fusion.is_left_protein_forward()),
                ("right", fusion.m_right_gene, fusion.m_right_gp, fusion.m_right_ref, fusion.m_right_ref_ext, fusion.m_right_pos,
                 fusion.m_right_is_exon, fusion.m_right_exon_or_intron_id, fusion.is_right_protein_forward())):
            w(f"\t\t\t\"{side}\":{{\n")
            w(f"\t\t\t\t\"gene_name\":\"{gene.m_name}\",\n")
            w(f"\t\t\t\t\"gene_chr\":\"{gene.m_chr}\",\n")
            w(f"\t\t\t\t\"position\":{gene.gene_pos_2_chr_pos(gp.position)},\n")
            w(f"\t\t\t\t\"reference\":\"{ref.decode()}\",\n")
            w(f"\t\t\t\t\"ref_ext\":\"{ext.decode()}\",\n")
            w(f"\t\t\t\t\"pos_str\":\"{pos}\",\n")
            w(f"\t\t\t\t\"exon_or_intron\":\"{'exon' if is_exon else 'intron'}\",\n")
            w(f"\t\t\t\t\"exon_or_intron_id\":{eid},\n")
            w(f"\t\t\t\t\"strand\":\"{'forward' if fwd else 'reversed'}\"\n")
            w("\t\t\t}, \n")
        w(f"\t\t\t\"unique\":{fusion.m_unique},\n")
        w("\t\t\t\"reads\":[\n")
        for m, me in enumerate(fusion.m_matches):
            w("\t\t\t\t{\n")
            w(f"\t\t\t\t\t\"break\":{me.m_read_break},\n")
            w(f"\t\t\t\t\t\"strand\":\"{'reversed' if me.m_reversed else 'forward'}\",\n")
            w(f"\t\t\t\t\t\"seq\":\"{me.m_read.m_seq.decode()}\",\n")
            w(f"\t\t\t\t\t\"qual\":\"{me.m_read.m_quality.decode()}\"\n")
            w("\t\t\t\t}")
            if m != len(fusion.m_matches) - 1:
                w(",")
            w("\n")
        w("\t\t\t]\n")
        w("\t\t}")
    w("\n\t}\n}\n\n")
    return "".join(o).encode()
