cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/r2x_pytest.log | cut -c1-220
HB2_SMALL_ILP=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hky85 or nuc or states or c1_" > gpurun_out/r2x_pytest_ilp.log 2>&1; echo pytest_ilp_rc=$?; tail -2 gpurun_out/r2x_pytest_ilp.log | cut -c1-220
rm -f gpurun_out/r2x_small*.json
timeout 300 python tools/bench_small.py >> gpurun_out/r2x_small.json 2>> gpurun_out/r2x_small.err
HB2_SMALL_ILP=1 timeout 300 python tools/bench_small.py >> gpurun_out/r2x_small_ilp1.json 2>/dev/null
HB2_SMALL_ILP=2 timeout 300 python tools/bench_small.py >> gpurun_out/r2x_small_ilp2.json 2>/dev/null
timeout 300 python tools/bench_small.py --taxa 256 --sites 120000 >> gpurun_out/r2x_small_120k.json 2>/dev/null
HB2_SMALL_ILP=1 timeout 300 python tools/bench_small.py --taxa 256 --sites 120000 >> gpurun_out/r2x_small_120k_ilp1.json 2>/dev/null
grep -ho '"workload": "[a-z0-9_]*"\|"pruning": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2x_small.json gpurun_out/r2x_small_ilp1.json gpurun_out/r2x_small_ilp2.json gpurun_out/r2x_small_120k.json gpurun_out/r2x_small_120k_ilp1.json | paste - - -
echo done
