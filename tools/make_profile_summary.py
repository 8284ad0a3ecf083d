#!/usr/bin/env python3
"""profiles/<tag>_summary.md + profiles/screen_traffic.json from one ncu launch list (--metrics gpu__time_duration.sum)
and one `ncu --set full` report of the six mapping kernels.
usage: make_profile_summary.py <tag> <launches.csv> <kernels.ncu-rep> "<workload text>" """
import collections, csv, json, os, subprocess, sys
tag, launches, rep, workload = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(launches, errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[14]) / 1e6
tot = sum(v[1] for v in agg.values())
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
raw = list(csv.reader(out.splitlines()))
hdr, units, krows = raw[0], raw[1], raw[2:]
def short(n):
    n = n.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    return n if len(n) < 40 else n[:12] + "…" + n[-22:]
names = [short(r[hdr.index("Kernel Name")]) for r in krows]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
def val(r, h):
    i = hdr.index(h)
    v = float(r[i].replace(",", "")); u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u)
    return v, u, scale
md = [f"# {tag}: ncu evidence of the mapping kernels", "", f"Workload: {workload}; `--clock-control none`.", "",
      f"## Launch list (`ncu --metrics gpu__time_duration.sum`, profiles/{os.path.basename(launches)}).  Cold-cache and serialised: compare SHARES.", "",
      "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k[:70]}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
md += ["", f"## `ncu --set full` (one launch each, gpurun_out/{os.path.basename(rep)}, not committed: > 10 MB)", "",
       "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
for h in want:
    if h not in hdr: continue
    i = hdr.index(h)
    md.append(f"| {h} [{units[i]}] | " + " | ".join(r[i] for r in krows) + " |")
# derived: achieved DRAM bandwidth of each launch against the measured copy peak (MEASURED_PEAKS.json)
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6538.3
gbs = []
for r in krows:
    rd, _, s1 = val(r, "dram__bytes_read.sum"); wr, _, s2 = val(r, "dram__bytes_write.sum"); t, _, st = val(r, "gpu__time_duration.sum")
    gbs.append((rd * s1 + wr * s2) / (t * st / 1e3) / 1e9)
md.append("| **derived: DRAM read+write [GB/s]** | " + " | ".join(f"{g:.0f}" for g in gbs) + " |")
md.append(f"| **derived: fraction of the measured HBM peak ({peak:.1f} GB/s)** | " + " | ".join(f"{g / peak:.2f}" for g in gbs) + " |")
reading = os.path.join(ROOT, "profiles", f"{tag}_reading.md")
if os.path.exists(reading):
    md += ["", open(reading).read().rstrip()]
kern = {}
for r, n in zip(krows, names):
    base = [b for b in ("k_prep", "k_seed", "k_diag", "k_scan", "k_exact", "k_verify") if b in r[hdr.index("Kernel Name")]]
    if not base: continue
    rd, _, s1 = val(r, "dram__bytes_read.sum"); wr, _, s2 = val(r, "dram__bytes_write.sum"); t, _, st = val(r, "gpu__time_duration.sum")
    kern[base[0]] = {"dram_bytes": int(rd * s1 + wr * s2), "ncu_ms": t * st,
                     "issue_active_pct": float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                     "l1tex_pct": float(r[hdr.index("l1tex__throughput.avg.pct_of_peak_sustained_active")]),
                     "l2_hit_pct": float(r[hdr.index("lts__t_sector_hit_rate.pct")]),
                     "warps_active_pct": float(r[hdr.index("sm__warps_active.avg.pct_of_peak_sustained_active")])}
screen = sum(kern[k]["dram_bytes"] for k in ("k_prep", "k_seed", "k_diag", "k_scan") if k in kern)
json.dump({"kernel": "k_prep+k_seed+k_diag+k_scan (screen v4)", "workload": workload, "dram_bytes_per_launch": screen,
           "kernels": kern, "source": f"profiles/{tag}_summary.md"}, open(os.path.join(ROOT, "profiles", "screen_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[-30:])); print(json.dumps(kern, indent=1))
