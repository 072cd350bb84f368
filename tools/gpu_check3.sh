#!/bin/bash
TAG=${1:-r01r}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 600 python tools/bench_small.py --states 4 --taxa 256 --sites 200000 --iters 10 > gpurun_out/${TAG}_small_d4.json 2> gpurun_out/${TAG}_small_d4.err; echo "small d4 rc=$?"; cat gpurun_out/${TAG}_small_d4.json; tail -2 gpurun_out/${TAG}_small_d4.err
HB2_SMALL_WALK=0 timeout 600 python tools/bench_small.py --states 4 --taxa 256 --sites 200000 --iters 10 > gpurun_out/${TAG}_small_d4_level.json 2> /dev/null; echo "small d4 level rc=$?"; cat gpurun_out/${TAG}_small_d4_level.json
timeout 600 python tools/bench_small.py --states 20 --taxa 128 --sites 20000 --classes 4 --iters 10 > gpurun_out/${TAG}_small_d20.json 2> gpurun_out/${TAG}_small_d20.err; echo "small d20 rc=$?"; cat gpurun_out/${TAG}_small_d20.json; tail -2 gpurun_out/${TAG}_small_d20.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_small_walk -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_small_d4 \
    python tools/bench_small.py --states 4 --taxa 256 --sites 200000 --iters 2 > gpurun_out/${TAG}_ncu_small.log 2>&1; echo "ncu small rc=$?"
