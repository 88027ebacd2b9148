timeout 100 python -m pytest tests -m gpu -x -q -k "decode" 2>&1 | tail -2
for e in 4 8; do
echo "== ICAP_DECODE_UB=$e"
ICAP_DECODE_UB=$e timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:round(v,1) for k,v in d['extra'].items() if 'ms_per' in k})
    else: print(l.rstrip()[-300:])
"
done
