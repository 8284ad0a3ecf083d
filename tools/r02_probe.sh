#!/bin/bash
# round-2 probe (developer tool): host topology of the GPU box, pinned H2D ceiling, panel-size sweep of the screen.
# usage (under gpurun): bash tools/r02_probe.sh
out=gpurun_out/r02_probe
mkdir -p $out
{
  echo "== nproc"; nproc
  echo "== lscpu"; lscpu | head -40
  echo "== numa nodes"; ls -d /sys/devices/system/node/node* 2>/dev/null
  for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist 2>/dev/null)"; done
  echo "== free"; free -g
  echo "== nvidia-smi topo"; nvidia-smi topo -m
  echo "== gpu pci numa"; for d in /sys/bus/pci/devices/*; do
      if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
  echo "== affinity"; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
  echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null
  echo "== pcie link"; nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
} > $out/topology.txt 2>&1
python tools/h2d_peak.py > $out/h2d_peak_1gpu.txt 2>&1
for s in 1 2 4 8; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --panel-scale $s > $out/panel_scale_$s.json 2> $out/panel_scale_$s.err
done
python - <<'PY'
import json
for s in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/r02_probe/panel_scale_{s}.json"))
        pk = d["roofline"]["per_kernel"]
        print(s, "pairs/s %.0fM step %.3f ms" % (d["value"] / 1e6, d["ms_per_step"]),
              {k: round(v["ms"], 3) for k, v in pk.items()}, "keys", d["config"]["index"]["keys"],
              "surv", d["survivors_per_step"], "matches", d["matches_per_step"])
    except Exception as e:
        print(s, "failed", e)
PY
