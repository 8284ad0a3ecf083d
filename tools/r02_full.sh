#!/bin/bash
# developer tool (under gpurun): what the driver runs at round end — GPU tests, smoke, both bench arms — plus the ncu evidence
tag=${1:-r02g}
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
(time python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err) 2>&1 | grep real
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2>/dev/null
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print("value %.1fM ms %.3f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity", {}).get("identical")))
print({k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["kernel_ms"], d["roofline"]["exact_verify_ms"])
e = d["e2e"]
print({k: e[k] for k in ("value", "h2d_bytes_per_step", "packed_upload", "pack_threads", "ms_per_call_min", "ms_per_call_median", "ms_per_call_median_ascii_upload")}, e["parity"] and e["parity"]["identical"])
for k, c in d["configs"].items():
    print(k, c.get("value"), c.get("unit"), c.get("ms_per_step"), (c.get("parity") or {}).get("identical"), c.get("roofline", {}).get("frac"))
r = json.load(open("$out/${tag}_bench_ref.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
CMD="python bench.py --steps 2 --warmup 3 --legs main --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_(prep|seed|diag|scan|exact|verify)$' -s 18 -c 6 -f -o $out/${tag}_kernels $CMD > $out/${tag}_ncu.log 2>&1
ls -la $out/${tag}_*
# the device inflate of the .fq.gz leg: its launches alone
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_bgzf --csv --log-file $out/${tag}_inflate_launches.csv python bench.py --legs fastq > /dev/null 2>&1
grep k_bgzf $out/${tag}_inflate_launches.csv | cut -d, -f5,14- | tail -3
