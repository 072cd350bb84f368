#!/bin/bash
# multi-GPU validation: bench at N GPUs (strong scaling over pattern shards, one ncclAllReduce per evaluation)
N=${1:-2}
TAG=${2:-r01m}
mkdir -p gpurun_out
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err; echo "n$n rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_n$n.json; tail -3 gpurun_out/${TAG}_bench_n$n.err
  fi
done
