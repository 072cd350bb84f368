#!/bin/bash
mkdir -p gpurun_out
for sh in 0/2 1/2 0/8; do
HB2_DEBUG=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --emulate-shard $sh > gpurun_out/dbg_shard.json 2> gpurun_out/dbg_shard.err; echo "shard $sh rc=$?"; cut -c1-200 gpurun_out/dbg_shard.json; grep -v "^\[hb2\] walk:" gpurun_out/dbg_shard.err | tail -5 | cut -c1-300; grep "walk:" gpurun_out/dbg_shard.err | sort | uniq -c | tail -3
done
