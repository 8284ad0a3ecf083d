#!/bin/bash
# e2e (gf_map_pairs from pinned host memory): pipeline chunk size, second copy stream, zero-copy qualities
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('e2e %.1fM pairs/s  h2d %.2f GB -> %.1f GB/s' % (d['e2e']['value']/1e6, d['e2e']['h2d_bytes_per_step']/1e9, d['e2e']['h2d_bytes_per_step']/1e9*d['e2e']['value']/1e7))"; }
echo -n "default             "; run
echo -n "GF_ZEROCOPY_QUAL=0  "; GF_ZEROCOPY_QUAL=0 run
echo -n "GF_CHUNK_MB=96      "; GF_CHUNK_MB=96 run
