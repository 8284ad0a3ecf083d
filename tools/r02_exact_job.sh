#!/bin/bash
# developer tool (under gpurun): k_exact changes — GPU tests, main leg, repeat-rich panel
tag=${1:-r02o}
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parity or noisy or repeat or edge" 2>&1 | tail -2
for rf in 0 0.3; do
  python bench.py --steps 10 --warmup 3 --legs main --no-e2e --repeat-frac $rf > $out/${tag}_rf$rf.json 2> $out/${tag}_rf$rf.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_rf$rf.json"))
print("repeat $rf: value %.1fM ms %.3f parity %s survivors %d" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity", {}).get("identical"), d["survivors_per_step"]), {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, round(d["roofline"]["exact_verify_ms"], 3))
PY
done
