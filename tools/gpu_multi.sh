#!/bin/bash
N=${1:-2}
TAG=${2:-r01m}
mkdir -p gpurun_out
HB2_DEBUG=1 NCCL_DEBUG=WARN timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "n$N rc=$?"; cut -c1-500 gpurun_out/${TAG}_bench_n$N.json; grep -E "EngineError|NCCL WARN|failed" gpurun_out/${TAG}_bench_n$N.err | head -5 | cut -c1-700
