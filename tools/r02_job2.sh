#!/bin/bash
# developer tool (under gpurun): all GPU tests, main bench leg with full-size parity, matcher leg
tag=${1:-r02c}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --legs main --steps 20 --warmup 5 --no-e2e > $out/${tag}_main.json 2> $out/${tag}_main.err
python - <<PY
import json
d = json.load(open("$out/${tag}_main.json"))
print("value %.1fM ms %.3f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity", {}).get("identical")))
print({k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["kernel_ms"], d["roofline"]["exact_verify_ms"])
PY
python bench.py --legs matcher --steps 3 > $out/${tag}_matcher.json 2> $out/${tag}_matcher.err
python - <<PY
import json
c = json.load(open("$out/${tag}_matcher.json"))["configs"]["matcher_pass"]
print({k: c[k] for k in ("ms_per_pass_e2e", "ms_scan_kernels", "h2d_gbs", "kernel_launches")}, c["roofline"]["frac"], c["roofline"]["kernel_ms"], c.get("parity", {}).get("identical"))
PY
