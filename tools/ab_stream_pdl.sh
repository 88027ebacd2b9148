for pr in 256 128 0; do
echo "== ICAP_TMA_L2_PROMOTION=$pr"
ICAP_TMA_L2_PROMOTION=$pr timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'gemm_ms', d['roofline']['gemm_ms_per_step'])
    else: print(l.rstrip()[-300:])
"
done
