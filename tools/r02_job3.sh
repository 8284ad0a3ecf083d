#!/bin/bash
# developer tool (under gpurun): all GPU tests, main bench leg incl. e2e (packed upload) with full-size parity
tag=${1:-r02f}
out=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --legs main --steps 20 --warmup 5 > $out/${tag}_main.json 2> $out/${tag}_main.err
tail -3 $out/${tag}_main.err
python - <<PY
import json
d = json.load(open("$out/${tag}_main.json"))
print("value %.1fM ms %.3f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity", {}).get("identical")))
print({k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["kernel_ms"], d["roofline"]["exact_verify_ms"])
e = d["e2e"]
print({k: e[k] for k in ("value", "h2d_bytes_per_step", "packed_upload", "pack_threads", "ms_per_call_min", "ms_per_call_median", "ms_per_call_median_ascii_upload")}, e["parity"] and e["parity"]["identical"])
PY
