for fa in 0 1; do
echo "== ICAP_DECODE_FUSED_APPEND=$fa"
ICAP_DECODE_STREAMS=1 ICAP_DECODE_FUSED_APPEND=$fa timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:round(v,1) for k,v in d['extra'].items() if 'ms_per' in k})
    else: print(l.rstrip()[-300:])
"
done
