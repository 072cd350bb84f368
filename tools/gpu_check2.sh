#!/bin/bash
# walk-kernel bring-up: tests with walk on, bench walk on/off
TAG=${1:-r01d}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_walk.json 2> gpurun_out/${TAG}_bench_walk.err; echo "bench walk rc=$?"; cat gpurun_out/${TAG}_bench_walk.json; tail -3 gpurun_out/${TAG}_bench_walk.err
HB2_TC_WALK=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_level.json 2> gpurun_out/${TAG}_bench_level.err; echo "bench level rc=$?"; cat gpurun_out/${TAG}_bench_level.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_walk \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
