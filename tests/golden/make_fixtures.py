#!/usr/bin/env python3
"""Regenerates tests/golden/* from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, the committed outputs travel instead).

Outputs
  kat.json           known-answer vectors lifted from the reference's own #[test] functions
                     (each entry cites file:line)
  testdata/          R1.fq R2.fq tinyref.fa fusions.csv — BASELINE config 1 inputs (data files, verbatim)
  cancer_genes.tsv   gene table derived from testdata/cancer.csv: name, chr, start, end, reversed, n_exons
                     (reversed per src/core/gene.rs:98-107) — the shape the synthetic panel is rebased from
"""
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def strings_in(path, lo, hi):
    lines = open(os.path.join(REF, path)).read().split("\n")[lo - 1:hi]
    return re.findall(r'"([^"]*)"', "\n".join(lines))


def main():
    kat = {}
    # fast_merge assert — src/core/read.rs:450-486
    s = strings_in("src/core/read.rs", 450, 480)
    kat["fast_merge"] = {
        "cite": "src/core/read.rs:450-486",
        "r1_name": s[0], "r1_seq": s[1], "r1_qual": s[3],
        "r2_name": s[4], "r2_seq": s[5], "r2_qual": s[7],
        "merged_seq": s[8],
    }
    # merged fixture of testdata pair #1 — src/core/indexer.rs:1059
    line = open(os.path.join(REF, "src/core/indexer.rs")).read().split("\n")[1058]
    m = re.search(r'm_name: "([^"]*)".*m_str: "([^"]*)".*m_quality: "([^"]*)"', line)
    kat["merged_fixture"] = {"cite": "src/core/indexer.rs:1059", "name": m.group(1), "seq": m.group(2),
                             "qual": m.group(3)}
    # edit distance [0, 1, 90] — src/core/edit_distance.rs:221-261
    s = strings_in("src/core/edit_distance.rs", 222, 233)
    kat["edit_distance"] = {"cite": "src/core/edit_distance.rs:221-261", "a": s[0:3], "b": s[3:6],
                            "expect": [0, 1, 90]}
    # reverse complement — src/core/sequence.rs:66-70
    s = strings_in("src/core/sequence.rs", 66, 70)
    kat["reverse_complement"] = {"cite": "src/core/sequence.rs:66-70",
                                 "pairs": [[s[1], s[0]], [s[3], s[2]]]}
    # gp_to_i64 round trip table — src/core/indexer.rs:982-983
    lines = open(os.path.join(REF, "src/core/indexer.rs")).read().split("\n")
    contigs = [int(x) for x in re.findall(r"-?\d+", lines[981].split("=")[1])]
    positions = [int(x) for x in re.findall(r"-?\d+", lines[982].split("=")[1])]
    kat["gp_roundtrip"] = {"cite": "src/core/indexer.rs:982-983", "contigs": contigs, "positions": positions}
    # tinyref contigs — src/core/fasta_reader.rs:237-238
    s = strings_in("src/core/fasta_reader.rs", 232, 245)
    kat["tinyref"] = {"cite": "src/core/fasta_reader.rs:232-279", "strings": s}
    # Gene::pos2str strings of testdata/fusions.csv — src/core/fusion.rs:116-141 (the test compares against these literals)
    lines = open(os.path.join(REF, "src/core/fusion.rs")).read().split("\n")[115:141]
    cases, gene = [], None
    for ln in lines:
        g = re.search(r'm_name == "([A-Z0-9]+)"', ln)
        if g:
            gene = g.group(1)
        c = re.search(r'pos2str\((-?\d+)\)\.unwrap\(\) != "([^"]*)"', ln)
        if c:
            cases.append([gene, int(c.group(1)), c.group(2)])
    kat["pos2str"] = {"cite": "src/core/fusion.rs:116-141 (over testdata/fusions.csv)", "cases": cases}
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    td = os.path.join(HERE, "testdata")
    os.makedirs(td, exist_ok=True)
    for f in ("R1.fq", "R2.fq", "tinyref.fa", "fusions.csv"):
        shutil.copy(os.path.join(REF, "testdata", f), os.path.join(td, f))

    # gene table of cancer.csv (parse rules: src/core/fusion.rs:23-91, src/core/gene.rs:40-42,98-107)
    rows = []
    cur = None
    for line in open(os.path.join(REF, "testdata", "cancer.csv")):
        line = line.strip()
        sp = line.split(",")
        if len(sp) < 2 or sp[0].startswith("#"):
            continue
        if sp[0].startswith(">"):
            if cur:
                rows.append(cur)
            name = sp[0][1:]
            chrom, rng = sp[1].split(":")
            a, b = rng.split("-")
            cur = {"name": name, "chr": chrom, "start": int(a), "end": int(b), "exons": []}
            continue
        if len(sp) >= 3:
            cur["exons"].append((int(sp[0]), int(sp[1]), int(sp[2])))
    if cur:
        rows.append(cur)
    with open(os.path.join(HERE, "cancer_genes.tsv"), "w") as f:
        f.write("#name\tchr\tstart\tend\treversed\tn_exons\n")
        for r in rows:
            ex = r["exons"]
            rev = int(len(ex) > 1 and ex[0][1] > ex[1][1])
            f.write(f"{r['name']}\t{r['chr']}\t{r['start']}\t{r['end']}\t{rev}\t{len(ex)}\n")
    print("genes:", len(rows), "bases:", sum(r["end"] - r["start"] for r in rows))


if __name__ == "__main__":
    main()
