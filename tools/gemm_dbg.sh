python tools/gemm_bench.py --bn 256,pair,auto 2>&1
