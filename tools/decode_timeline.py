"""Device timeline of ONE replay of the graphed decode (CUPTI through torch.profiler): start, duration and the gap to
the previous kernel's end for every kernel of one decode step -- shows what the per-step time is made of (kernel
durations vs. launch gaps vs. overlap).
    python tools/decode_timeline.py [--beam 5] [--batch 512] [--step 10] [--out gpurun_out/decode_timeline.csv]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402  (synthetic inputs only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--step", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/decode_timeline.csv")
    args = ap.parse_args()
    pkg = icap_loader.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    kw = dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="bench",
              dropout=0.2)
    model = pkg.Transformer(device=dev, **kw).to(dev).eval()
    f, p, _ = O.synthetic_batch(args.batch, 36, 2048, 84, 22, 10000, seed=4321)
    f, p = f.to(dev), p.to(dev)
    gd = pkg.GraphedDecode(model, args.batch, 36, args.beam)
    for _ in range(3):
        gd.run(f, p)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        gd.run(f, p)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    evs = [e for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    rows = []
    prev_end = t0
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        rows.append((e.name, s - t0, en - s, s - prev_end))
        prev_end = max(prev_end, en)
    total = prev_end - t0
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        fh.write("idx,kernel,start_us,dur_us,gap_after_prev_end_us\n")
        for i, (n, s, d, g) in enumerate(rows):
            fh.write(f"{i},\"{n[:90]}\",{s:.2f},{d:.2f},{g:.2f}\n")
    print(f"{len(rows)} kernels, {total:.1f} us from first start to last end (profiled replay)")
    # one decode step: from one step-start kernel (decode_embed_ln; embed_fwd in the unfused build) to the next
    marks = [i for i, r in enumerate(rows) if "decode_embed_ln" in r[0]]
    if not marks:
        marks = [i for i, r in enumerate(rows) if "embed_fwd" in r[0]]
    if len(marks) > args.step + 1:
        a, b = marks[args.step], marks[args.step + 1]
        seg = rows[a:b]
        print(f"-- step {args.step + 1}: {len(seg)} kernels, {seg[-1][1] + seg[-1][2] - seg[0][1]:.1f} us; "
              f"sum of durations {sum(r[2] for r in seg):.1f} us, sum of positive gaps {sum(max(0, r[3]) for r in seg):.1f} us")
        for n, s, d, g in seg:
            short = n.split("<")[0].split("::")[-1][:28]
            print(f"   {short:28s} start {s - seg[0][1]:8.2f}  dur {d:7.2f}  gap {g:6.2f}")
    agg = {}
    for n, s, d, g in rows:
        k = n.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1; a[1] += d; a[2] += max(0.0, g)
    print("-- whole decode, by kernel: launches, total us, total gap-before us")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {a[0]:5d} {a[1]:9.1f} {a[2]:8.1f}  {k}")


if __name__ == "__main__":
    main()
