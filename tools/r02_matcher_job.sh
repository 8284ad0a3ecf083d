#!/bin/bash
# developer tool (under gpurun): Matcher tests, the matcher bench leg, and an ncu capture of k_ref_scan on the resident reference
tag=${1:-r02b}
out=gpurun_out
timeout 600 python -m pytest tests/test_matcher.py -m gpu -x -q 2>&1 | tail -3
python bench.py --legs matcher --steps 3 > $out/${tag}_matcher.json 2> $out/${tag}_matcher.err
python - <<PY
import json
c = json.load(open("$out/${tag}_matcher.json"))["configs"]["matcher_pass"]
print({k: c[k] for k in ("ms_per_pass_e2e", "ms_pass_device_span", "ms_scan_kernels", "h2d_gbs", "kernel_launches")}, c["roofline"]["frac"], c["roofline"]["kernel_ms"], c.get("parity", {}).get("identical"))
PY
# the resident-reference launches come after the warm-up pass and the three host passes (17 launches each)
ncu --set full --clock-control none --import-source on -k regex:'^k_ref_scan' -s 70 -c 1 -f -o $out/${tag}_refscan python bench.py --legs matcher --steps 3 --no-cpu-baseline > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_refscan.ncu-rep --page details 2>/dev/null | grep -E "Duration|Registers Per|Achieved Occupancy|Issue Slots Busy|DRAM Throughput|Executed Ipc Active|No Eligible|Grid Size|Memory Throughput|L2 Hit|Executed Instructions" | head -12

