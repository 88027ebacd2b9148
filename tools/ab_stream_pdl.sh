timeout 200 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "micro_batched or full_size_train" 2>&1 | tail -3
for n in 1 2 4; do
echo "== ICAP_MICRO_BATCHES=$n"
ICAP_MICRO_BATCHES=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'loss', d['final_loss'], 'e2e', d['e2e']['value'])
    else: print(l.rstrip()[-300:])
"
done
