#!/bin/bash
# developer tool (under gpurun): the repeat-rich panel (30 % of the bases in 2-5-copy blocks): bench line + ncu of k_exact / k_verify
tag=${1:-r02n}
out=gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --legs main --no-e2e --repeat-frac 0.3"
$CMD > $out/${tag}_repeat.json 2> $out/${tag}_repeat.err
python - <<PY
import json
d = json.load(open("$out/${tag}_repeat.json"))
print("value %.1fM ms %.3f parity %s survivors %d matches %d" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity", {}).get("identical"), d["survivors_per_step"], d["matches_per_step"]))
print({k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["kernel_ms"], d["roofline"]["exact_verify_ms"])
PY
ncu --set full --clock-control none --import-source on -k regex:'^k_(exact|verify)' -s 8 -c 2 -f -o $out/${tag}_exact $CMD --no-cpu-baseline > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_exact.ncu-rep --page details 2>/dev/null | grep -E "^\s+(Duration|Registers Per|Achieved Occupancy|Issue Slots Busy|Executed Ipc Active|No Eligible|Grid Size|Executed Instructions|L2 Hit|Block Limit)|k_exact|k_verify" | head -30
