"""In-kernel timeline of a chain of DEPENDENT decode-step GEMMs (M = 2560 rows): every launch's CTA 0 records
%globaltimer stamps (icap_debug_trace) -- kernel entry, prologue done, grid dependency resolved, first operand stage
landed, last MMA issued, accumulator ready, epilogue stores issued, exit -- so the fixed cost of a launch can be split
into its parts.  Compares the persistent kernel (ICAP_GEMM_SMALL=0) with the small-footprint kernel (=1).
    python tools/gemm_trace.py [--rows 2560] [--layers 6]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

pkg = icap_loader.load()
N = pkg._native
BF16 = N.BF16
NAMES = ["entry", "prologue", "dep_ok", "loads_issued", "first_full", "mma_issued", "acc_ready", "stores_issued", "exit"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2560)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--static", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    M, d, F = args.rows, 512, 2048
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, d, device=dev, generator=g).bfloat16()
    wqkv = (torch.randn(3 * d, d, device=dev, generator=g) * 0.04).bfloat16()
    wo = (torch.randn(d, d, device=dev, generator=g) * 0.04).bfloat16()
    w1 = (torch.randn(F, d, device=dev, generator=g) * 0.04).bfloat16()
    w2 = (torch.randn(d, F, device=dev, generator=g) * 0.02).bfloat16()
    b1 = torch.zeros(F, device=dev)
    qkv = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
    o = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    h = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    y = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    flag = N.EPI_B_STATIC if args.static else 0
    N.call("icap_set_pdl", 1)

    def gemm(a, lda, w, n, k, c, bias=None, epi=0):
        N.call("icap_gemm", BF16, 1, 1, M, n, k, a.data_ptr(), lda, w.data_ptr(), k, c.data_ptr(), n, BF16,
               bias.data_ptr() if bias is not None else None, epi | flag, None, 0, 0, 1, torch.cuda.current_stream().cuda_stream)

    def chain():
        a = x
        for _ in range(args.layers):
            gemm(a, d, wqkv, 3 * d, d, qkv)
            gemm(qkv, 3 * d, wo, d, d, o)
            gemm(o, d, w1, F, d, h, bias=b1, epi=1)
            gemm(h, F, w2, d, F, y)
            a = y
    shapes = ["qkv 1536x512", "proj 512x512", "ffn1 2048x512", "ffn2 512x2048"]
    nl = 4 * args.layers
    for mode in ("0", "1"):
        os.environ["ICAP_GEMM_SMALL"] = mode
        N.call("icap_reload_env")
        N.call("icap_debug_trace", None, 0)
        for _ in range(3):
            chain()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            chain()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"== ICAP_GEMM_SMALL={mode}: {nl} dependent launches, {us:.1f} us per chain = {us / nl:.2f} us per launch")
        # traced replay
        buf = torch.zeros(nl, 16, dtype=torch.int64, device=dev)
        N.call("icap_debug_trace", buf.data_ptr(), nl)
        gt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gt):
            chain()
        N.call("icap_debug_trace", None, 0)
        gt.replay()
        gt.replay()
        torch.cuda.synchronize()
        t = buf.cpu()
        base = int(t[0, 0])
        print("launch  shape            " + " ".join(f"{n:>13s}" for n in NAMES) + "   (ns after the first launch's entry)")
        for i in range(nl):
            row = [int(v) - base if int(v) else -1 for v in t[i, :9]]
            print(f"{i:4d}    {shapes[i % 4]:16s} " + " ".join(f"{v:13d}" for v in row))
    os.environ.pop("ICAP_GEMM_SMALL", None)


if __name__ == "__main__":
    main()
