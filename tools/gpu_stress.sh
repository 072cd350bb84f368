#!/bin/bash
# usage: gpu_stress.sh TAG lib1 lib2 ...   (lib = file name under hyphy_b200/_alt, or "default")
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset HB2_LIB; else export HB2_LIB=$PWD/hyphy_b200/_alt/libhyphy_b200_$v.so; fi
  timeout 600 python tools/stress_determinism.py 8 > gpurun_out/${TAG}_stress_$v.log 2>&1; echo "$v rc=$?"; tail -5 gpurun_out/${TAG}_stress_$v.log | cut -c1-300
done
