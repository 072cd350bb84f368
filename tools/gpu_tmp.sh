cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/r2y_pytest_multigpu.log 2>&1; echo pytest_multi_rc=$?; tail -4 gpurun_out/r2y_pytest_multigpu.log | cut -c1-250
bash tools/gpu_multi.sh 2 r2y
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 3 --fp64 > gpurun_out/r2y_bench_n2_fp64.json 2> gpurun_out/r2y_bench_n2_fp64.err; echo fp64_n2_rc=$?; cut -c1-300 gpurun_out/r2y_bench_n2_fp64.json
echo done
