#!/bin/bash
# Round-end style validation: smoke, full GPU test-suite, default bench (with cpu_baseline), reference arm, ncu evidence.
TAG=${1:-r01z}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 300 python tools/stress_determinism.py 3 > gpurun_out/${TAG}_stress.log 2>&1; echo "stress rc=$?"; tail -2 gpurun_out/${TAG}_stress.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json | cut -c1-2500; tail -2 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host --no-c5 --no-small > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune64_tc_walk -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_walk \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host --no-c5 --no-small > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune64_lanes -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_lanes \
    python bench.py --fp64 --steps 3 --warmup 3 --no-cpu-baseline --no-host --no-c5 --no-small > gpurun_out/${TAG}_ncu_full_fp64.log 2>&1; echo "ncu full fp64 rc=$?"
timeout 300 python bench.py --fp64 --steps 20 --warmup 3 --no-cpu-baseline --no-host --no-small > gpurun_out/${TAG}_bench_fp64.json 2>/dev/null; echo "bench fp64 rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_fp64.json
timeout 600 python tools/host_e2e.py > gpurun_out/${TAG}_host_e2e.jsonl 2> gpurun_out/${TAG}_host_e2e.err; echo "host e2e rc=$?"; cut -c1-100 gpurun_out/${TAG}_host_e2e.jsonl
timeout 300 python tools/bench_branch_cache.py 40 > gpurun_out/${TAG}_branch_cache.json 2> gpurun_out/${TAG}_branch_cache.err; echo "branch cache rc=$?"; cut -c1-900 gpurun_out/${TAG}_branch_cache.json
