for mb in 48; do
echo "== ICAP_DP_BUCKET_MB=$mb"
ICAP_DP_BUCKET_MB=$mb timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$mb bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'loss', d['final_loss'])
    elif 'warn' in l.lower() or 'error' in l.lower() or 'Traceback' in l: print(l.rstrip()[-400:])
"
done
