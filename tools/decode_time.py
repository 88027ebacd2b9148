"""Timing stability probe: several fresh GraphedDecode objects (beam 5, 512 images) in one process."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402

pkg = icap_loader.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
kw = dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="bench", dropout=0.2)
model = pkg.Transformer(device=dev, **kw).to(dev).eval()
f, p, _ = O.synthetic_batch(512, 36, 2048, 84, 22, 10000, seed=4321)
f, p = f.to(dev), p.to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
keep = []
for trial in range(4):
    gd = pkg.GraphedDecode(model, 512, 36, 5)
    for _ in range(2):
        gd.run(f, p)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record(); gd.run(f, p); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    print("trial", trial, ts, flush=True)
    if trial % 2 == 0:
        keep.append(torch.empty(300 << 20, dtype=torch.uint8, device=dev))   # perturb the allocator between trials
    del gd
