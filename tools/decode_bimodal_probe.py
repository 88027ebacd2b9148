"""Why is beam-5 decode slower when its graph is captured after an eager training step?  Times a fresh GraphedDecode
(1) in a clean process, (2) after an eager train step whose activations were all kept alive and then freed (what
bench.py's roofline section does), (3) after torch.cuda.empty_cache()."""
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


def timed(tag):
    gd = pkg.GraphedDecode(model, 512, 36, 5)
    for _ in range(2):
        gd.run(f, p)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record(); gd.run(f, p); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    st = torch.cuda.memory_stats()
    print(tag, ts, "reserved MB", torch.cuda.memory_reserved() >> 20, "segments", st["segment.all.current"], flush=True)
    del gd


timed("clean          ")
ft, pt, ct = O.synthetic_batch(256, 36, 2048, 84, 22, 10000, seed=1)
ft, pt, ct = ft.to(dev), pt.to(dev), ct.to(dev)
eng = model._engine()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
model.train()
if which in ("all", "fwd"):
    with torch.no_grad():
        model(ft, pt, ct)
    torch.cuda.synchronize()
    model.eval()
    timed("after eager fwd (no_grad, train mode)")
    model.train()
if which in ("all", "fb"):
    eng.forward_backward(ft, pt, ct)
    torch.cuda.synchronize()
    model.eval()
    timed("after eager fwd+bwd")
    model.train()
if which in ("all", "step"):
    eng.train_step(ft, pt, ct, lr=5e-4)
    torch.cuda.synchronize()
    model.eval()
    timed("after eager step   ")
    eng.shadow_fresh = False
    eng.refresh_shadow()
    timed("after refresh_shadow")
    model.train()
gs = pkg.GraphedTrainStep(model.train(), 256, 36, 22, lr=5e-4)
gs.load(ft, pt, ct)
for _ in range(3):
    gs.step()
torch.cuda.synchronize()
model.eval()
timed("after graphstep")
model.train()
eng.train_step(ft, pt, ct, lr=5e-4)
torch.cuda.synchronize()
model.eval()
timed("after eager step 2 ")
