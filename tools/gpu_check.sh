#!/bin/bash
# one GPU round trip: parity tests, a short bench, and per-kernel time / instruction counts of one step
# usage (under gpurun): bash tools/gpu_check.sh <tag>
tag=${1:-chk}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}.json"))
print("pairs/s %.1fM  step %.3f ms  screen %.3f ms  frac %.3f  survivors %d  matches %d" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["survivors_per_step"], d["matches_per_step"]))
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'^k_(prep|seed|diag|scan|exact|verify)$' -c 8 --csv --log-file gpurun_out/${tag}_k.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_k.csv")) if len(r)>14 and r[0].isdigit()]
by={}
for r in rows:
    by.setdefault((r[0],r[4].split("(")[0][-30:]),{})[r[12]]=r[14]
for (i,k),m in by.items():
    print(k, " ".join("%s=%s"%(a.split("__")[-1][:22],b) for a,b in m.items()))
PY
