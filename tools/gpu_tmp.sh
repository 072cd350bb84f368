cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2z4_pytest.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/r2z4_pytest.log | cut -c1-220
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-host --no-small > gpurun_out/r2z4_bench.json 2> gpurun_out/r2z4_bench.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z4_bench.json').readline())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stage_ms'), d.get('c5',{}).get('value'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2z4_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host --no-c5 --no-small > /dev/null 2>&1; python tools/ncu_summary.py launches gpurun_out/r2z4_launches.csv gpurun_out/r2z4_launches.md; cat gpurun_out/r2z4_launches.md | tail -10
echo done
