#!/bin/bash
TAG=${1:-r01h}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
HB2_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_split.json 2> gpurun_out/${TAG}_bench_split.err; echo "bench split rc=$?"; sort gpurun_out/${TAG}_bench_split.err | uniq -c | tail -3
HB2_WALK_SPLIT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_nosplit.json 2> gpurun_out/${TAG}_bench_nosplit.err; echo "bench nosplit rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --fp64 > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench fp64 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_walk \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expm64_dmma -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_expm \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
python - <<'PY'
import json,glob,sys
tag=sys.argv[1] if len(sys.argv)>1 else ''
for f in sorted(glob.glob('gpurun_out/*_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['lnL'])
    except Exception as e: print(f, 'ERR', e)
PY
