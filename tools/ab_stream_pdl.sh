for pdl in 0 1; do
echo "== PDL=$pdl"
ICAP_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'loss', d['final_loss'], 'gemm_ms', d['roofline']['gemm_ms_per_step']); print({k:round(v,1) for k,v in d['extra'].items()})
    else: print(l.rstrip()[-300:])
"
done
