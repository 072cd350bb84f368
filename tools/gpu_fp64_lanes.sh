#!/bin/bash
# body of a gpurun call: parity + A/B timing of the fp64 lanes kernel (prune64_lanes_kernel).  usage: tools/gpu_fp64_lanes.sh TAG
TAG=${1:-r2m}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo pytest_rc=$?; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-220
B="--steps 20 --warmup 3 --fp64 --no-cpu-baseline --no-host --no-small"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo bench_rc=$?
HB2_FP64_WALK=0 timeout 300 python bench.py $B --no-c5 > gpurun_out/${TAG}_bench_fp64_levels.json 2>/dev/null
HB2_WALK_SPLIT_NODES=0 timeout 300 python bench.py $B --no-c5 > gpurun_out/${TAG}_bench_fp64_nosplit.json 2>/dev/null
HB2_WALK_LANES=4 timeout 300 python bench.py $B --no-c5 > gpurun_out/${TAG}_bench_fp64_k4.json 2>/dev/null
HB2_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 1 --fp64 --no-cpu-baseline --no-host --no-small --no-c5 2>&1 | grep "fp64 lanes\|schedule" | sort | uniq -c | head -8
timeout 600 python tools/host_e2e.py > gpurun_out/${TAG}_host_e2e.jsonl 2> gpurun_out/${TAG}_host_e2e.err; echo host_rc=$?
grep -h '"pruning"' gpurun_out/${TAG}_bench_fp64*.json | cut -c1-400
echo done
