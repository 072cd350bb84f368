#!/bin/bash
# partition churn of the workflow tests through the patched host: how many engine partitions a run creates and where their
# lifetime goes.  usage: tools/gpu_churn.sh TAG test...
TAG=$1; shift
cd "${GRAFT_REPO_ROOT:-/root/repo}"
HB2_REGRESS_KEEP=1 timeout 900 python host/regress.py check "$@" > gpurun_out/${TAG}_regress.jsonl 2>&1
python - <<PY
import json
for l in open('gpurun_out/${TAG}_regress.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['test'], d['ok'], d['got']['seconds'], 'cpu', d['expected'].get('seconds_cpu'))
PY
for t in "$@"; do
  f=gpurun_out/regress_$(echo $t | tr / _).fp64.txt
  echo "$f: partitions $(grep -c 'on device' $f)"
  grep "destroyed after" $f | sed -e 's/.*destroyed after \([0-9]*\) evaluations.*lifetime \([0-9.]*\) s of which \([0-9.]*\) s in matrix hand-over and \([0-9.]*\) s in hb2_evaluate.*/\1 \2 \3 \4/' | awk '{n++; ev+=$1; life+=$2; ho+=$3; evt+=$4} END {print "destroyed", n, "evaluations", ev, "lifetime_s", life, "handover_s", ho, "evaluate_s", evt}'
  grep "on device" $f | sed -e 's/.*: \([0-9]*\) patterns x.*/\1/' | sort -n | uniq -c | sort -rn | head -4
done
