#!/bin/bash
# How much of the PRECISE FFN kernel is weight streaming?  CATSEG_DBG_NOSTREAM=1 fetches only the first ring images (results
# are WRONG; timing only).  Also: the row-shift probe of the tcgen05 operand reads.
for m in 0 1; do
  echo "== CATSEG_DBG_NOSTREAM=$m"
  CATSEG_DBG_NOSTREAM=$m python bench.py --steps 3 --warmup 2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()})"
done
tools/probes/bin/umma_probe 26 2>&1 | grep PROBE
