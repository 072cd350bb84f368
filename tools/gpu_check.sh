#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench + fp64 microbenchmark + ncu launch list + one full capture.
# usage (from the repo root, on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_nvsmi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb_fp64 tools/microbench_fp64.cu && /tmp/mb_fp64 > gpurun_out/${TAG}_mb_fp64.txt 2>&1; cat gpurun_out/${TAG}_mb_fp64.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune64 -s 40 -c 2 -f -o gpurun_out/${TAG}_prof_prune64 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expm64 -s 4 -c 1 -f -o gpurun_out/${TAG}_prof_expm64 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
