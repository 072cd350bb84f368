#!/bin/bash
TAG=${1:-r01g}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
HB2_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_walk.json 2> gpurun_out/${TAG}_bench_walk.err; echo "bench walk rc=$?"; sort gpurun_out/${TAG}_bench_walk.err | uniq -c | tail -4
HB2_DEBUG=1 HB2_WALK_CTAS_PER_SM=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_walk2.json 2> gpurun_out/${TAG}_bench_walk2.err; echo "bench walk 2/SM rc=$?"; sort gpurun_out/${TAG}_bench_walk2.err | uniq -c | tail -4
HB2_TC_WALK=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_level.json 2> gpurun_out/${TAG}_bench_level.err; echo "bench level rc=$?"
HB2_WALK_CTAS_PER_SM=2 timeout 600 python -m pytest tests -m gpu -q -x -k "golden or partial or full_size or tensor_path" > gpurun_out/${TAG}_pytest_gpu_2sm.log 2>&1; echo "pytest 2/SM rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu_2sm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_walk \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/'+''+''+''+'_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['lnL'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
