for v in 1 2 3 6; do
echo "== ICAP_LN_COLS_WAVES=$v"
ICAP_LN_COLS_WAVES=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'])
    else: print(l.rstrip()[-300:])
"
done
