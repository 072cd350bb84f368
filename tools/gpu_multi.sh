#!/bin/bash
# multi-GPU bench on N GPUs of one box: class groups x pattern shards (default) and patterns only, tight timeouts
N=${1:-2}
TAG=${2:-r01m}
mkdir -p gpurun_out
run() {   # name, extra flags
  HB2_DEBUG=${HB2_DEBUG:-0} NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 \
     bench.py --gpus $N --steps 20 --warmup 3 $2 > gpurun_out/${TAG}_bench_n${N}_$1.json 2> gpurun_out/${TAG}_bench_n${N}_$1.err; echo "n$N $1 rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/${TAG}_bench_n${N}_$1.json') if l.startswith('{')][-1]
    print('$1', d.get('layout'), round(d['value'],1), 'evals/s e2e', round(d['e2e']['value'],1), 'lnL', repr(d.get('lnL')), 'c5', (d.get('c5') or {}).get('value'))
except Exception as e:
    print('$1 ERR', e)
PY
  grep -E "EngineError|NCCL WARN|failed|Error" gpurun_out/${TAG}_bench_n${N}_$1.err | head -5 | cut -c1-600
}
run cg ""
if [ "$N" -ge 4 ]; then run cg2 "--class-groups 2"; fi
if [ "${ALSO_PATTERNS_ONLY:-0}" = 1 ]; then run pat "--no-class-groups"; fi
