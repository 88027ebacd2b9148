"""The decode step's classifier GEMM (rows x V x d, bf16, bias) alone: plain epilogue against ICAP_EPI_ROWSTATS (per-row,
per-128-column largest / second largest / sum-exp for icap_beam_select), back-to-back launches timed with CUDA events,
plus CTA 0's in-kernel stamps (icap_debug_trace) of one launch.
    python tools/classifier_probe.py [--rows 2560] [--vocab 10000] [--d 512]"""
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
    ap.add_argument("--vocab", type=int, default=10000)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--beam", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    M, V, d = args.rows, args.vocab, args.d
    ldl = (V + 7) // 8 * 8
    P = 2 * ((V + 255) // 256)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, d, device=dev, generator=g).bfloat16()
    w = (torch.randn(V, d, device=dev, generator=g) * 0.04).bfloat16()
    bias = torch.randn(V, device=dev, generator=g) * 0.1
    logits = torch.empty(M, ldl, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(M, 4 * P, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    N.call("icap_set_pdl", 1)

    def run(epi):
        N.call("icap_gemm", BF16, 1, 1, M, V, d, x.data_ptr(), d, w.data_ptr(), d, logits.data_ptr(), ldl, BF16, bias.data_ptr(),
               epi | N.EPI_B_STATIC, stats.data_ptr() if epi == N.EPI_ROWSTATS else None, 4 * P if epi == N.EPI_ROWSTATS else 0,
               0, 1, st)

    for name, epi in (("plain", N.EPI_NONE), ("rowstats", N.EPI_ROWSTATS)):
        N.call("icap_debug_trace", None, 0)
        for _ in range(5):
            run(epi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            run(epi)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        tf = 2.0 * M * V * d / us * 1e-6
        print(f"== {name}: {us:.2f} us per launch back to back = {tf:.0f} TFLOP/s")
        buf = torch.zeros(4, 16, dtype=torch.int64, device=dev)
        N.call("icap_debug_trace", buf.data_ptr(), 4)
        for _ in range(4):
            run(epi)
        torch.cuda.synchronize()
        N.call("icap_debug_trace", None, 0)
        t = buf.cpu()[3]
        base = int(t[0])
        print("   CTA 0 of the 4th launch, us after entry: " +
              "  ".join(f"{n} {(int(t[i]) - base) / 1e3:.2f}" for i, n in enumerate(NAMES) if int(t[i])))
    # the consumer: icap_beam_select on the logits just produced, scanning them itself / fed with the statistics
    k = args.beam
    B = M // k
    prev = torch.rand(B, k, device=dev, generator=g) * 1e-3
    osc = torch.empty(B, k, device=dev)
    opar = torch.empty(B, k, dtype=torch.int32, device=dev)
    otok = torch.empty(B, k, dtype=torch.int32, device=dev)
    run(N.EPI_ROWSTATS)
    for name, use in (("scan", False), ("stats", True)):
        def sel():
            N.call("icap_beam_select", BF16, B, k, V, logits.data_ptr(), ldl, prev.data_ptr(), k, osc.data_ptr(), opar.data_ptr(),
                   otok.data_ptr(), None, 0, stats.data_ptr() if use else None, 4 * P if use else 0, st)
        for _ in range(5):
            sel()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            sel()
        e1.record()
        torch.cuda.synchronize()
        print(f"== beam_select ({name}): {e0.elapsed_time(e1) * 1e3 / args.iters:.2f} us per launch back to back")


if __name__ == "__main__":
    main()
