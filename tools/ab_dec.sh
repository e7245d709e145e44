timeout 600 python -m pytest tests/test_gpu_precise.py -m gpu -q -s -x --timeout 300 -k "precise_stage and decoder" 2>&1 | grep -E "max-abs|passed|failed" | head -5
for ns in 0 200 1000; do echo "== backoff $ns"; CATSEG_DEC_BACKOFF_NS=$ns python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['roofline']['stage_ms_per_step']['decoder'])"; done
CATSEG_PHASE_TIMING=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra 2>&1 >/dev/null | grep -E "band_conv" | tail -5
