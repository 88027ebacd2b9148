#!/bin/bash
# A/B of the graphed beam-5 decode (512 images) with the batch cut into 1/2/4 parallel graph branches.
for n in 1 2 4; do
  echo "== ICAP_DECODE_STREAMS=$n"
  ICAP_DECODE_STREAMS=$n timeout 300 python tools/decode_time.py 2>&1 | tail -2
done
