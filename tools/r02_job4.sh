#!/bin/bash
# developer tool (under gpurun): GPU tests + main leg (device-timed, full-size parity) + per-kernel DRAM bytes from ncu
tag=${1:-r02i}
out=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --legs main --steps 20 --warmup 5 > $out/${tag}_main.json 2> $out/${tag}_main.err
python - <<PY
import json
d = json.load(open("$out/${tag}_main.json"))
print("value %.1fM ms %.3f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity", {}).get("identical")))
print({k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["kernel_ms"], d["roofline"]["exact_verify_ms"])
e = d["e2e"]
print({k: e[k] for k in ("value", "h2d_bytes_per_step", "packed_upload", "pack_threads", "ms_per_call_min", "ms_per_call_median", "ms_per_call_median_ascii_upload", "ms_host_pack_per_step")}, e["parity"] and e["parity"]["identical"])
PY
CMD="python bench.py --steps 2 --warmup 3 --legs main --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'^k_(prep|seed|diag|scan)$' -s 12 -c 4 --csv --log-file $out/${tag}_dram.csv $CMD > /dev/null 2>&1
python - <<PY
import csv, re
rows = [r for r in csv.reader(open("$out/${tag}_dram.csv", errors="replace")) if len(r) > 14 and r[0].isdigit()]
k = {}
for r in rows:
    k.setdefault(re.search(r"(k_\w+)", r[4]).group(1), {})[r[12]] = float(r[14].replace(",", ""))
tot = 0
for n, v in k.items():
    gb = (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / 1e9; tot += gb
    print(n, "ms %.3f" % (v["gpu__time_duration.sum"] / 1e6), "read %.2f GB write %.2f GB" % (v["dram__bytes_read.sum"] / 1e9, v["dram__bytes_write.sum"] / 1e9), "L2 hit %.1f" % v["lts__t_sector_hit_rate.pct"])
print("screen DRAM total %.2f GB" % tot)
PY
