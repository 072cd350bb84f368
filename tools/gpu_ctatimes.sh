#!/bin/bash
TAG=${1:-ct}
mkdir -p gpurun_out
HB2_DEBUG=1 timeout 200 python tools/trace_walk.py 2 > gpurun_out/trace_2.log 2>&1; echo "rc=$?"; cp gpurun_out/walk_trace_cta2.txt gpurun_out/${TAG}_walk_trace_cta2.txt
