"""Micro-benchmark of icap_gemm (bf16 tcgen05) on the GEMM shapes of one training step of model A
(batch 256): CUDA-event time per launch in a back-to-back loop, for both tile widths.
    python tools/gemm_bench.py [--iters 50]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

pkg = icap_loader.load()
N = pkg._native
F32, BF16 = N.F32, N.BF16

# (name, a_k, b_k, M, N, K, c_dtype, bias, epi, accumulate, split_k, count per step)
SHAPES = [
    ("enc qkv fwd", 1, 1, 9216, 1536, 512, BF16, 0, 0, 0, 1, 6),
    ("enc joint fwd", 1, 1, 9216, 512, 512, BF16, 0, 0, 0, 1, 6),
    ("enc ffn1 fwd", 1, 1, 9216, 2048, 512, BF16, 1, 1, 0, 1, 6),
    ("enc ffn2 fwd", 1, 1, 9216, 512, 2048, BF16, 1, 0, 0, 1, 6),
    ("enc embed fwd", 1, 1, 9216, 512, 2136, BF16, 0, 0, 0, 1, 1),
    ("cross kv fwd", 1, 1, 9216, 1024, 512, BF16, 0, 0, 0, 1, 6),
    ("dec qkv fwd", 1, 1, 5376, 1536, 512, BF16, 0, 0, 0, 1, 6),
    ("dec d x d fwd", 1, 1, 5376, 512, 512, BF16, 0, 0, 0, 1, 19),
    ("dec ffn1 fwd", 1, 1, 5376, 2048, 512, BF16, 1, 1, 0, 1, 6),
    ("dec ffn2 fwd", 1, 1, 5376, 512, 2048, BF16, 1, 0, 0, 1, 6),
    ("classifier fwd", 1, 1, 5376, 10000, 512, BF16, 1, 0, 0, 1, 1),
    ("enc qkv dgrad", 1, 0, 9216, 512, 1536, BF16, 0, 0, 0, 1, 6),
    ("enc d x d dgrad", 1, 0, 9216, 512, 512, BF16, 0, 0, 0, 1, 6),
    ("enc ffn2 dgrad+mask", 1, 0, 9216, 2048, 512, BF16, 0, 2, 0, 1, 6),
    ("enc ffn1 dgrad", 1, 0, 9216, 512, 2048, BF16, 0, 0, 0, 1, 6),
    ("cross kv dgrad (acc)", 1, 0, 9216, 512, 1024, BF16, 0, 0, 1, 1, 6),
    ("dec ffn2 dgrad+mask", 1, 0, 5376, 2048, 512, BF16, 0, 2, 0, 1, 6),
    ("classifier dgrad", 1, 0, 5376, 512, 10000, BF16, 0, 0, 0, 1, 1),
    ("enc qkv wgrad", 0, 0, 1536, 512, 9216, F32, 0, 0, 1, 0, 6),
    ("enc d x d wgrad", 0, 0, 512, 512, 9216, F32, 0, 0, 1, 0, 6),
    ("enc ffn1 wgrad", 0, 0, 2048, 512, 9216, F32, 0, 0, 1, 0, 6),
    ("enc ffn2 wgrad", 0, 0, 512, 2048, 9216, F32, 0, 0, 1, 0, 6),
    ("embed wgrad", 0, 0, 512, 2136, 9216, F32, 0, 0, 1, 0, 1),
    ("dec d x d wgrad", 0, 0, 512, 512, 5376, F32, 0, 0, 1, 0, 19),
    ("dec ffn1 wgrad", 0, 0, 2048, 512, 5376, F32, 0, 0, 1, 0, 6),
    ("classifier wgrad", 0, 0, 10000, 512, 5376, F32, 0, 0, 1, 0, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--bn", default="256,pair,auto")
    ap.add_argument("--only", default="", help="substring filter on the shape name")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (for ncu)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = {}
    print(f"{'shape':24s} {'M':>6s} {'N':>6s} {'K':>6s} " + " ".join(f"{'us@' + b:>10s} {'TF/s':>7s}" for b in args.bn.split(",")))
    for name, ak, bk, M, Nn, K, cdt, bias, epi, acc, split, cnt in SHAPES:
        if args.only and args.only not in name:
            continue
        A = torch.randn((M, K) if ak else (K, M), device=dev).bfloat16()
        B = torch.randn((Nn, K) if bk else (K, Nn), device=dev).bfloat16()
        C = torch.zeros(M, Nn, device=dev, dtype=torch.float32 if cdt == F32 else torch.bfloat16)
        bias_t = torch.randn(Nn, device=dev) if bias else None
        aux = torch.randn(M, Nn, device=dev).to(C.dtype) if epi == 2 else None
        row = f"{name:24s} {M:6d} {Nn:6d} {K:6d} "
        for bn in args.bn.split(","):
            if bn == "auto":
                os.environ.pop("ICAP_GEMM_BN", None)
            else:
                os.environ["ICAP_GEMM_BN"] = bn

            def go():
                N.call("icap_gemm", BF16, ak, bk, M, Nn, K, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], C.data_ptr(),
                       Nn, cdt, bias_t.data_ptr() if bias else None, epi, aux.data_ptr() if aux is not None else None, Nn,
                       acc, split, torch.cuda.current_stream().cuda_stream)
            for _ in range(5):
                go()
            torch.cuda.synchronize()
            if args.eager:
                row += "  (eager: not timed) "
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # CUDA graph of `iters` back-to-back launches: no host launch overhead in the measurement
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(args.iters):
                    go()
            gr.replay()
            torch.cuda.synchronize()
            flush.zero_()
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / args.iters
            tf = 2.0 * M * Nn * K / (us * 1e-6) / 1e12
            row += f"{us:10.2f} {tf:7.1f} "
            tot[bn] = tot.get(bn, 0.0) + us * cnt
        print(row)
    print("per-step GEMM time (us), launches weighted by count:", {k: round(v, 1) for k, v in tot.items()})


if __name__ == "__main__":
    main()
