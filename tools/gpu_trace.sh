#!/bin/bash
# kernel-iteration loop on one GPU: determinism stress, tc parity tests, bench line, per-step trace of one walk CTA
TAG=${1:-r01i}
mkdir -p gpurun_out
timeout 600 python tools/stress_determinism.py 4 > gpurun_out/${TAG}_stress.log 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/${TAG}_stress.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --emulate-shard 0/8 > gpurun_out/${TAG}_bench_s8.json 2> gpurun_out/${TAG}_bench_s8.err; echo "bench shard rc=$?"
for c in 0 1 2 3; do HB2_DEBUG=1 timeout 200 python tools/trace_walk.py $c > gpurun_out/trace_$c.log 2>&1; echo "cta $c rc=$?"; cp gpurun_out/walk_trace_cta$c.txt gpurun_out/${TAG}_walk_trace_cta$c.txt; done
python - <<PY
import json
for f in ['gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_bench_s8.json']:
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]; print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['lnL'])
    except Exception as e: print(f, 'ERR', e)
PY
