#!/bin/bash
# A/B of the data-parallel knobs at N GPUs (default 2): bucket size, SMs reserved for NCCL during the backward, NCCL CTA cap.
N=${1:-2}
run() {
  echo "== $*"
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ms_per_step %.3f  samples/s %.0f  e2e %.0f  dp_parity %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['extra'].get('dp_parity_rel_err')))
    elif 'error' in l.lower() or 'Traceback' in l: print(l.rstrip()[-300:])
"
}
run ICAP_DP_GRAD_DTYPE=fp32
run ICAP_DP_GRAD_DTYPE=bf16
run ICAP_DP_GRAD_DTYPE=fp32 ICAP_DP_BUCKET_MB=16
run ICAP_DP_GRAD_DTYPE=fp32 ICAP_DP_BUCKET_MB=16 ICAP_DP_RESERVE_SMS=8 NCCL_MAX_CTAS=8
