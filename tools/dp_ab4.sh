for n in 4 2; do for peer in 0 1 0 1; do ICAP_DP_PEER=$peer timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 80 --warmup 5 --no-cpu-baseline --no-decode 2>/dev/null | grep "^{" | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('N', j['n_gpus'], 'PEER=$peer', round(j['ms_per_step'],4), round(j['value']), round(j['e2e']['value']))"; done; done
python bench.py --steps 80 --warmup 5 --no-cpu-baseline --no-decode 2>/dev/null | grep "^{" | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('N', j['n_gpus'], round(j['ms_per_step'],4), round(j['value']), round(j['e2e']['value']))"
