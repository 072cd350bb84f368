#!/bin/bash
# multi-GPU validation: bench at N = 1 and N (strong scaling over pattern shards, one ncclAllReduce per evaluation)
N=${1:-2}
TAG=${2:-r01m}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err; echo "n$n rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_n$n.json; tail -3 gpurun_out/${TAG}_bench_n$n.err
  fi
done
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json
