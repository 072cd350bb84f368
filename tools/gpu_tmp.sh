cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2s_pytest.log | cut -c1-220
HB2_LANES_WARPS=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp64" > gpurun_out/r2s_pytest_nw8.log 2>&1; echo pytest_nw8_rc=$?; tail -2 gpurun_out/r2s_pytest_nw8.log | cut -c1-220
B="--steps 20 --warmup 3 --fp64 --no-cpu-baseline --no-host --no-small"
timeout 300 python bench.py $B > gpurun_out/r2s_bench_fp64.json 2> gpurun_out/r2s_bench_fp64.err; echo bench_rc=$?
HB2_LANES_WARPS=4 timeout 300 python bench.py $B > gpurun_out/r2s_bench_fp64_nw4.json 2>/dev/null
HB2_LANES_WARPS=8 timeout 300 python bench.py $B --no-c5 > gpurun_out/r2s_bench_fp64_nw8.json 2>/dev/null
HB2_WALK_SPLIT_NODES=0 timeout 300 python bench.py $B --no-c5 > gpurun_out/r2s_bench_fp64_nosplit.json 2>/dev/null
grep -h '"ms_per_step"' gpurun_out/r2s_bench_fp64*.json | cut -c100-250
grep -ho '"c5": {[^}]*}' gpurun_out/r2s_bench_fp64*.json | cut -c1-300
HB2_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 3 --fp64 --no-cpu-baseline --no-host --no-small 2>&1 >/dev/null | grep "lanes" | sort | uniq -c | head
timeout 600 python tools/host_e2e.py > gpurun_out/r2s_host_e2e.jsonl 2> gpurun_out/r2s_host_e2e.err; echo host_rc=$?
cut -c1-120 gpurun_out/r2s_host_e2e.jsonl
echo done
