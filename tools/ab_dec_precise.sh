#!/bin/bash
# A/B of the PRECISE decoder shapes: CATSEG_DECS_NARROW bit i = stage D(i+1) uses the 2-CTA/SM shape
for m in 0 2 4 8 16 30; do
  echo "== CATSEG_DECS_NARROW=$m"
  CATSEG_DECS_NARROW=$m python bench.py --steps 3 --warmup 2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()})"
done
echo "== phase timing (decoder stages), default shapes"
CATSEG_PHASE_TIMING=1 python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline --batch 4 2>&1 >/dev/null | grep band_conv | tail -5
