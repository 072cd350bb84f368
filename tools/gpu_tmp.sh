cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo pytest_rc=$?; tail -4 gpurun_out/r2u_pytest.log | cut -c1-220
for st in 20 16 29; do timeout 300 python tools/bench_small.py --states $st --taxa 128 --sites 20000 --classes 4 >> gpurun_out/r2u_small.json 2>> gpurun_out/r2u_small.err; done; echo small_rc=$?; cut -c1-420 gpurun_out/r2u_small.json
for st in 20 16 29; do HB2_SMALL_DMMA=0 timeout 300 python tools/bench_small.py --states $st --taxa 128 --sites 20000 --classes 4 >> gpurun_out/r2u_small_nodmma.json 2>/dev/null; done; cut -c1-420 gpurun_out/r2u_small_nodmma.json
timeout 300 python tools/bench_branch_cache.py 40 > gpurun_out/r2u_branch_cache.json 2> gpurun_out/r2u_branch_cache.err; echo bc_rc=$?; cut -c1-700 gpurun_out/r2u_branch_cache.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_small_dmma -s 4 -c 1 -f -o gpurun_out/r2u_prof_small_dmma python tools/bench_small.py --states 20 --taxa 128 --sites 20000 --classes 4 > gpurun_out/r2u_ncu_small.log 2>&1; echo ncu_rc=$?
echo done
