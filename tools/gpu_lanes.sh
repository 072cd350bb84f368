#!/bin/bash
# lane-count sweep of the walk kernel on emulated pattern shards (one GPU)
TAG=${1:-lanes}
mkdir -p gpurun_out
for cfg in "0/2 8" "0/4 8" "0/4 12" "0/4 16" "0/8 8" "0/8 16" "0/8 24" "0/8 32"; do
  set -- $cfg
  HB2_WALK_LANES=$2 timeout 300 python bench.py --emulate-shard $1 --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('shard $1 lanes $2:', round(d['value'],1), 'evals/s', d['roofline']['stage_ms'], 'e2e', round(d['e2e']['value'],1), 'lnL', d['lnL'])
"
done | tee gpurun_out/${TAG}_sweep.txt
