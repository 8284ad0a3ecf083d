#!/bin/bash
# developer tool: where does the screen leave L2?  panel size x1/2/4/8 and a repeat-rich panel (30 % of the bases in 2-5-copy
# blocks): device-timed bench line + ncu L2 hit rate / DRAM bytes / duration of the four screen kernels of one step.
out=gpurun_out/r02_panel
mkdir -p $out
for cfg in "1 0" "2 0" "4 0" "8 0" "1 0.3"; do
  set -- $cfg
  tag=s${1}_r${2}
  CMD="python bench.py --steps 10 --warmup 3 --legs main --no-cpu-baseline --no-e2e --panel-scale $1 --repeat-frac $2"
  $CMD > $out/$tag.json 2> $out/$tag.err &&
  ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:'^k_(prep|seed|diag|scan)$' -s 12 -c 4 --csv --log-file $out/$tag.ncu.csv $CMD > /dev/null 2>&1
done
python - <<'PY'
import csv, json, glob, os, re
for f in sorted(glob.glob("gpurun_out/r02_panel/*.json")):
    d = json.load(open(f)); tag = os.path.basename(f)[:-5]
    line = {"panel_scale": d["config"]["panel_scale"], "repeat_frac": d["config"]["repeat_frac"], "keys": d["index"]["keys"],
            "index_bytes": d["index"]["device_bytes"], "pairs_per_s": d["value"], "ms_per_step": d["ms_per_step"],
            "survivors": d["survivors_per_step"], "matches": d["matches_per_step"],
            "kernel_ms": {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, "clocks": d["clocks"]}
    try:
        rows = [r for r in csv.reader(open(f[:-5] + ".ncu.csv", errors="replace")) if len(r) > 14 and r[0].isdigit()]
        k = {}
        for r in rows:
            name = re.search(r"(k_\w+)", r[4]).group(1)
            k.setdefault(name, {})[r[12]] = r[14] + " " + r[13]
        line["ncu"] = k
    except Exception as e:
        line["ncu"] = str(e)
    print(json.dumps(line))
PY
