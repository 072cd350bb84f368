#!/bin/bash
TAG=${1:-lt}
mkdir -p gpurun_out
for c in 0 1 2 3; do timeout 200 python tools/trace_walk.py $c > gpurun_out/trace_$c.log 2>&1; echo "cta $c rc=$?"; cp gpurun_out/walk_trace_cta$c.txt gpurun_out/${TAG}_walk_trace_cta$c.txt; done
