#!/bin/bash
# usage: tools/ab_env.sh VAR v1 v2 ...   -- bench line (value + stage times + parity) per value of an A/B environment switch
VAR=$1; shift
for v in "$@"; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --steps 5 --warmup 3 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()})
p=d.get('parity') or {}
print({k:p.get(k) for k in ('max_abs','argmax_raw','mask_equal')})"
done
