#!/bin/bash
TAG=${1:-r01i}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hyphy_b200/csrc -o /tmp/tcm tools/tc_mma_timing.cu && timeout 60 /tmp/tcm > gpurun_out/${TAG}_mma_timing.txt 2>&1; cat gpurun_out/${TAG}_mma_timing.txt
HB2_WALK_SPLIT=0 timeout 600 python -m pytest tests -m gpu -q -x -k "tc" > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -4 gpurun_out/${TAG}_pytest_tc.log
HB2_WALK_SPLIT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_nosplit.json 2> gpurun_out/${TAG}_bench_nosplit.err; echo "bench rc=$?"
for c in 2; do timeout 200 python tools/trace_walk.py $c > gpurun_out/trace_$c.log 2>&1; echo "cta $c rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r01i_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['lnL'])
    except Exception as e: print(f, 'ERR', e)
PY
