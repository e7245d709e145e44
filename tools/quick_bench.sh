#!/bin/bash
# usage: tools/quick_bench.sh [bench.py args]   -- one compact line: value, stage times, parity of the benchmarked precision
python bench.py --steps 5 --warmup 3 --no-extra "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get('parity') or {}
print(round(d['value'],1), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()}, p.get('max_abs'), p.get('argmax_raw'))"
