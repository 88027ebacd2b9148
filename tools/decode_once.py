"""One eager beam-5 decode of 512 images (model A, bf16) -- target for ncu launch lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402  (synthetic inputs only)

pkg = icap_loader.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
kw = dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="bench", dropout=0.2)
model = pkg.Transformer(device=dev, **kw).to(dev).eval()
eng = model._engine()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
f, p, _ = O.synthetic_batch(512, 36, 2048, 84, 22, 10000, seed=4321)
f, p = f.to(dev), p.to(dev)
with torch.no_grad():
    eng.decode(f, p, beam_size=k)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.decode(f, p, beam_size=k)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
