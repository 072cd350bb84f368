cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "protein_sized" > gpurun_out/r2w_pytest.log 2>&1; echo pytest_rc=$?; grep -n "AssertionError\|passed\|failed" gpurun_out/r2w_pytest.log | cut -c1-400
HB2_SMALL_DMMA=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "protein_sized" > gpurun_out/r2w_pytest_nodmma.log 2>&1; echo pytest_rc=$?; grep -n "AssertionError\|passed\|failed" gpurun_out/r2w_pytest_nodmma.log | cut -c1-400
echo done
