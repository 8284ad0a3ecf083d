#!/usr/bin/env python3
"""Aggregates an ncu report's per-instruction counters by source function / line (needs -lineinfo builds).
usage: ncu_by_function.py report.ncu-rep [n_units] [kernel-regex]   (n_units: divide instruction counts, e.g. pairs per launch)"""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kf = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kf, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) > 2 and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    num = lambda x: int(x) if x.isdigit() else 0
    a = agg[(cur, ln)]; a[0] += num(r[4]); a[1] += num(r[ie]); a[2] = r[1][:100]
ti = sum(v[1] for v in agg.values()) or 1; ts = sum(v[0] for v in agg.values()) or 1
print(f"total warp instructions {ti}  ({ti/units:.1f} per unit), samples {ts}")
funcs = {}
def func_of(path, ln):
    if path not in funcs:
        marks = []
        try:
            for i, l in enumerate(open(path, errors="replace").read().split("\n"), 1):
                if l.startswith("__device__") or l.startswith("__global__"):
                    nm = re.findall(r"(\w+)\(", l); marks.append((i, nm[0] if nm else "?"))
        except OSError: pass
        funcs[path] = marks
    name = os.path.basename(path)
    for a, n in funcs[path]:
        if a <= ln: name = os.path.basename(path) + ":" + n
    return name
byf = collections.defaultdict(lambda: [0, 0])
for (f, ln), v in agg.items():
    k = func_of(f, ln); byf[k][0] += v[0]; byf[k][1] += v[1]
print("--- by function")
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{100*v[1]/ti:5.1f}% inst ({v[1]/units:8.1f}/unit) {100*v[0]/ts:5.1f}% samples  {k}")
print("--- top lines")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{100*v[1]/ti:5.1f}% inst {100*v[0]/ts:5.1f}% smp  {os.path.basename(k[0])}:{k[1]}  {v[2]}")
