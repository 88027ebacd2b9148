timeout 300 python -m pytest tests -m gpu -x -q -k "linear_res_ln or decode or golden or odd_shapes" 2>&1 | tail -2
for v in 0 1; do
echo "== ICAP_FUSED_PROJ_LN=$v"
ICAP_FUSED_PROJ_LN=$v timeout 300 python tools/decode_time.py 2>&1 | head -2
done
