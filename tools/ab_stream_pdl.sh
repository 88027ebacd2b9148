timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do
echo "== ICAP_XKV_SIDE=$v"
ICAP_XKV_SIDE=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'loss', d['final_loss'])
    else: print(l.rstrip()[-300:])
"
done
