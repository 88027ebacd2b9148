#!/bin/bash
# Data-parallel A/B on N GPUs (default 8) through bench.py: gradient exchange back-ends and their knobs.
#   bash tools/scale8.sh [N] [set]      set = backends (default) | knobs
N=${1:-8}
SET=${2:-backends}
run() {
  echo "== $*"
  env "${@:2}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline $1 2>&1 | grep "^{" | tee -a gpurun_out/r2_scale${N}_lines.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); e=d['extra']
    print('  ms_per_step %.3f  samples/s %.0f  e2e %.0f  cache_e2e %s  dp_parity %s  beam5 %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], e.get('train_region_cache_e2e_samples_per_s'), e.get('dp_parity_rel_err'), e.get('beam5_captions_per_s')))
"
}
if [ "$SET" = backends ]; then
  run "--no-decode" ICAP_DP_PEER=0
  run "--no-decode" ICAP_DP_PEER=1 ICAP_DP_BUCKET_MB=16
  run "--no-decode" ICAP_DP_PEER=1 ICAP_DP_BUCKET_MB=32
  run "--workload modelC" ICAP_DP_PEER=0
else
  run "--no-decode" ICAP_DP_PEER_CTAS=148
  run "--no-decode" ICAP_DP_PEER_CTAS=64
  run "--no-decode" ICAP_DP_PEER_CTAS=32
  run "--no-decode" ICAP_DP_TAIL_MB=2
  run "--no-decode" ICAP_DP_BUCKET_MB=8 ICAP_DP_TAIL_MB=2
  run "--workload modelC" ICAP_DP_PEER_CTAS=148
fi
