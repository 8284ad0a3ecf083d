#!/bin/bash
# e2e (gf_map_pairs from pinned host memory) against the pipeline chunk size
for mb in 48 96 192 384 768; do
  echo -n "GF_CHUNK_MB=$mb  "
  GF_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('e2e %.1fM pairs/s  h2d %.2f GB' % (d['e2e']['value']/1e6, d['e2e']['h2d_bytes_per_step']/1e9))"
done
