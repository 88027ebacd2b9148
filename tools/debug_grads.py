"""GPU debug helper: per-parameter gradient error of the CUDA path vs the CPU oracle (model A, small batch)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402

pkg = icap_loader.load()
dev = torch.device("cuda:0")
kw = dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="x", dropout=0.0)
cfg = O.OracleConfig(**kw)
sd = O.init_state_dict(cfg, seed=0)
f, p, c = O.synthetic_batch(6, 36, 2048, 84, 22, 10000, seed=1234)
ref_loss, ref_grads = O.loss_and_grads(sd, cfg, f, p, c)
for precision in ("fp32", "bf16"):
    m = pkg.Transformer(device=dev, **kw)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    m.set_precision(precision)
    loss = m(f, p, c)["loss"]
    loss.backward()
    rows = []
    for name, q in m.named_parameters():
        r = ref_grads[name]
        g = q.grad.cpu()
        mx = float((g - r).abs().max() / (r.abs().max() + 1e-12))
        fro = float((g - r).norm() / (r.norm() + 1e-12))
        rows.append((mx, fro, name))
    rows.sort(reverse=True)
    print(f"== {precision}: loss {float(loss.detach()):.6f} ref {float(ref_loss):.6f}")
    for mx, fro, name in rows[:12]:
        print(f"  max {mx:.2e}  fro {fro:.2e}  {name}")
    name = rows[0][2]
    g = dict(m.named_parameters())[name].grad.cpu()
    d = (g - ref_grads[name]).abs()
    if d.dim() == 2:
        per_row = d.max(1).values
        top = torch.topk(per_row, min(5, per_row.numel()))
        print("  worst param rows:", [(int(i), f"{float(v):.2e}") for v, i in zip(top.values, top.indices)],
              "median row err", f"{float(per_row.median()):.2e}", "ref max", f"{float(ref_grads[name].abs().max()):.2e}")
