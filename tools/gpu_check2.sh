#!/bin/bash
# walk-kernel / DMMA expm bring-up: tests, bench variants, ncu
TAG=${1:-r01e}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
HB2_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_walk.json 2> gpurun_out/${TAG}_bench_walk.err; echo "bench walk rc=$?"; cat gpurun_out/${TAG}_bench_walk.json; sort gpurun_out/${TAG}_bench_walk.err | uniq -c | tail -5
HB2_EXPM_DFMA=1 HB2_TC_WALK=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_old.json 2> gpurun_out/${TAG}_bench_old.err; echo "bench old rc=$?"; cat gpurun_out/${TAG}_bench_old.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --fp64 > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench fp64 rc=$?"; cat gpurun_out/${TAG}_bench_fp64.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_walk \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expm64_dmma -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_expm \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
