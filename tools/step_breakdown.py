"""In-situ time breakdown of one training step (model A, batch 256) by C-ABI entry point.

Records every libicap call of one eager step (name + arguments), then re-launches (a) all of them and (b) each
entry point's calls alone from a CUDA graph and times the replays with CUDA events.  Same buffers, shapes and
launch order as the real step; activations stay allocated, so the replays are memory-safe.  Complements the ncu
launch list (whose per-kernel times are cold-cache: ncu flushes L2 between kernels).
    python tools/step_breakdown.py [--decode K]   (K: also break down a beam-K decode of 512 images)"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402  (synthetic inputs only)

pkg = icap_loader.load()
N = pkg._native


def replay_time(calls, reps=5):
    if not calls:
        return 0.0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for name, args in calls:
            N.call(name, *args[:-1], st)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def breakdown(title, calls):
    by = collections.OrderedDict()
    for c in calls:
        by.setdefault(c[0], []).append(c)
    total = replay_time(calls)
    print(f"## {title}: {len(calls)} launches, all replayed together: {total:.1f} us")
    rows = []
    for name, cs in by.items():
        rows.append((replay_time(cs), name, len(cs)))
    ssum = sum(r[0] for r in rows)
    print("| entry point | launches | us (replayed alone) | us / launch | share of sum |\n|---|---:|---:|---:|---:|")
    for t, name, n in sorted(rows, reverse=True):
        print(f"| `{name}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / ssum:.1f} % |")
    print(f"| sum | {len(calls)} | {ssum:.1f} | | |")
    return by


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decode", type=int, default=0)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    kw = dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="bench",
              dropout=0.2)
    model = pkg.Transformer(device=dev, **kw).to(dev).train()
    eng = model._engine()
    f, p, c = O.synthetic_batch(args.batch, 36, 2048, 84, 22, 10000, seed=1234)
    f, p, c = f.to(dev), p.to(dev), c.to(dev)
    for _ in range(2):
        eng.train_step(f, p, c)
    torch.cuda.synchronize()
    keep = []
    orig_new = eng.new

    def pinned_new(*a, **k):
        t = orig_new(*a, **k)
        keep.append(t)
        return t
    eng.new = pinned_new
    N.call_log = []
    eng.train_step(f, p, c)
    calls, N.call_log = N.call_log, None
    torch.cuda.synchronize()
    by = breakdown(f"train step, batch {args.batch}", calls)
    # GEMMs by role
    gemms = by.get("icap_gemm", [])
    roles = collections.OrderedDict()
    for cl in gemms:
        a = cl[1]
        role = "wgrad (A,B MN-major, fp32 C)" if (a[1], a[2]) == (0, 0) else ("dgrad (B MN-major)" if a[2] == 0 else "forward")
        roles.setdefault(role, []).append(cl)
    print("\n| GEMM role | launches | us | TFLOP/s |\n|---|---:|---:|---:|")
    for role, cs in roles.items():
        t = replay_time(cs)
        fl = sum(2.0 * x[1][3] * x[1][4] * x[1][5] for x in cs)
        print(f"| {role} | {len(cs)} | {t:.1f} | {fl / t / 1e6:.0f} |")
    eng.new = orig_new
    del keep

    if args.decode:
        model.eval()
        fd, pd, _ = O.synthetic_batch(512, 36, 2048, 84, 22, 10000, seed=4321)
        fd, pd = fd.to(dev), pd.to(dev)
        eng.decode(fd, pd, beam_size=args.decode)
        torch.cuda.synchronize()
        keep2 = []

        def pinned_new2(*a, **k):
            t = orig_new(*a, **k)
            keep2.append(t)
            return t
        eng.new = pinned_new2
        N.call_log = []
        out = eng.decode(fd, pd, beam_size=args.decode)
        calls, N.call_log = N.call_log, None
        torch.cuda.synchronize()
        print()
        breakdown(f"beam-{args.decode} decode, 512 images (torch fill/arange kernels not included)", calls)
        eng.new = orig_new


if __name__ == "__main__":
    main()
