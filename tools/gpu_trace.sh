#!/bin/bash
TAG=${1:-r01t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "tc" > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -4 gpurun_out/${TAG}_pytest_tc.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
for c in 2; do timeout 200 python tools/trace_walk.py $c > gpurun_out/trace_$c.log 2>&1; echo "cta $c rc=$?"; done
python - <<'PY'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/*_bench.json'))[-1:]:
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['lnL'])
    except Exception as e: print(f, 'ERR', e)
PY
