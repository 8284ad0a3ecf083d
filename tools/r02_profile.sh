#!/bin/bash
# round-2 profile job (developer tool): plain run, ncu launch list, ncu --set full of the six mapping kernels and of k_ref_scan.
# usage (under gpurun): bash tools/r02_profile.sh <tag>
tag=${1:-r02}
out=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --legs main --no-cpu-baseline --no-e2e"
$CMD > $out/${tag}_plain.json 2> $out/${tag}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_(prep|seed|diag|scan|exact|verify)$' -s 18 -c 6 -f -o $out/${tag}_kernels $CMD > $out/${tag}_ncu.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --legs matcher --no-cpu-baseline --matcher-mbases 512"
$CMD2 > $out/${tag}_matcher_plain.json 2> $out/${tag}_matcher_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'^k_ref_scan$' -s 12 -c 1 -f -o $out/${tag}_refscan $CMD2 > $out/${tag}_ncu2.log 2>&1
ls -la $out/${tag}_*
