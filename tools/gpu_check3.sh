#!/bin/bash
TAG=${1:-r01s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "other_state or golden or per_class or compiled or partial" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python tools/bench_small.py --states 4 --taxa 256 --sites 200000 --iters 10 > gpurun_out/${TAG}_small_d4.json 2> gpurun_out/${TAG}_small_d4.err; echo "small d4 rc=$?"; cat gpurun_out/${TAG}_small_d4.json; tail -2 gpurun_out/${TAG}_small_d4.err
timeout 600 python tools/bench_small.py --states 20 --taxa 128 --sites 20000 --classes 4 --iters 10 > gpurun_out/${TAG}_small_d20.json 2> gpurun_out/${TAG}_small_d20.err; echo "small d20 rc=$?"; cat gpurun_out/${TAG}_small_d20.json; tail -2 gpurun_out/${TAG}_small_d20.err
