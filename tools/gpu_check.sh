#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench (tensor + fp64) + ncu launch list + full captures.
# usage (from the repo root, on the GPU box): bash tools/gpu_check.sh [tag] [quick]
TAG=${1:-r01}
QUICK=${2:-}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hyphy_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu && timeout 120 /tmp/tc_probe > gpurun_out/${TAG}_tc_probe.txt 2>&1; cat gpurun_out/${TAG}_tc_probe.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 --fp64 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench fp64 rc=$?"; cat gpurun_out/${TAG}_bench_fp64.json
if [ -z "$QUICK" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune64 -s 60 -c 3 -f -o gpurun_out/${TAG}_prof_prune64 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expm64 -s 4 -c 1 -f -o gpurun_out/${TAG}_prof_expm64 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
fi
