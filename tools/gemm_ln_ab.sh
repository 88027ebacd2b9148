#!/bin/bash
# A/B of the fused projection+LayerNorm cluster kernel (ICAP_GEMM_LN): kernel parity, decode time, train step time
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm_ln" 2>&1 | tail -3
for cfg in "0 auto" "1 128" "1 256"; do
  set -- $cfg
  echo "== decode ICAP_GEMM_LN=$1 BN=$2"
  if [ "$2" = auto ]; then unset ICAP_GEMM_LN_BN; else export ICAP_GEMM_LN_BN=$2; fi
  ICAP_GEMM_LN=$1 timeout 200 python tools/decode_time.py 2>&1 | tail -1
done
for cfg in "0 auto" "2 auto" "2 256" "2 128"; do
  set -- $cfg
  echo "== train ICAP_GEMM_LN=$1 BN=$2"
  if [ "$2" = auto ]; then unset ICAP_GEMM_LN_BN; else export ICAP_GEMM_LN_BN=$2; fi
  ICAP_GEMM_LN=$1 timeout 200 python bench.py --no-decode --no-cpu-baseline --steps 30 2>/dev/null 
done
