"""Repro for tests/test_model_gpu.py::test_train_step_fused_auto_graph_equals_eager (fp32 mode): run the 5-batch
sequence eagerly and through the auto-captured graphs several times and print, per step, the relative difference of the
gradient buffer and of the Adam first moment between a run and the first eager run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402

pkg = icap_loader.load()
DEV = torch.device("cuda:0")
kw = dict(num_vocab=500, max_length=22, encode_dim_positions=84, encode_dim_features=256, output_name="x", dropout=0.0,
          encode_num_blocks=1, decode_num_blocks=1)
sd = O.init_state_dict(O.OracleConfig(**kw), seed=0)
batches = [O.synthetic_batch(16, 12, 256, 84, 22, 500, seed=20), O.synthetic_batch(16, 12, 256, 84, 22, 500, seed=21),
           O.synthetic_batch(9, 12, 256, 84, 22, 500, seed=22), O.synthetic_batch(16, 12, 256, 84, 22, 500, seed=23),
           O.synthetic_batch(9, 7, 256, 84, 22, 500, seed=24)]


def run(graph):
    os.environ["ICAP_TRAIN_GRAPH"] = graph
    m = pkg.Transformer(device=DEV, **kw)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    m.set_precision("fp32")
    eng = m._engine()
    out = []
    for f, p, c in batches:
        loss = float(m.train_step_fused(f, p, c, lr=5e-4, train_mode=False))
        torch.cuda.synchronize()
        out.append((loss, eng.g32[:eng.n_flat].clone(), eng.adam_m.clone()))
    return out


base = run("0")
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    for graph in ("0", "1"):
        cur = run(graph)
        row = []
        for (l0, g0, m0), (l1, g1, m1) in zip(base, cur):
            row.append(f"{abs(l0 - l1) / abs(l0):.0e}/{float((g0 - g1).norm() / g0.norm()):.0e}/{float((m0 - m1).norm() / m0.norm()):.0e}")
        print(f"trial {trial} graph={graph}  loss/grad/m rel diff per step: " + "  ".join(row), flush=True)

# ---- which tensors differ between the two trajectories?  (eager runs only)
os.environ["ICAP_TRAIN_GRAPH"] = "0"


def run2():
    m = pkg.Transformer(device=DEV, **kw)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    m.set_precision("fp32")
    eng = m._engine()
    snaps = []
    for f, p, c in batches[:2]:
        m.train_step_fused(f, p, c, lr=5e-4, train_mode=False)
        torch.cuda.synchronize()
        snaps.append((eng.g32[:eng.n_flat].clone(), eng.p32.clone()))
    return eng, snaps


eng0, ref = run2()
found = None
for i in range(20):
    e, s = run2()
    d = float((s[1][0] - ref[1][0]).norm() / ref[1][0].norm())
    if d > 1e-5:
        found = s
        break
if found is None:
    print("second trajectory not seen in 20 runs")
else:
    names = sorted(eng0.offsets.items(), key=lambda kv: kv[1])
    dW = (found[0][1] - ref[0][1]).abs()
    idx = torch.nonzero(dW > 1e-6).flatten()
    print("elements whose weight after step 1 differs by > 1e-6:", idx.numel(), idx[:20].tolist())
    for i in idx[:10].tolist():
        print("   idx", i, "w_ref", float(ref[0][1][i]), "w_alt", float(found[0][1][i]), "g1_ref", float(ref[0][0][i]), "g1_alt", float(found[0][0][i]))
    print("per-parameter differences between the two trajectories: |dW after step 1| max, rel grad diff at step 1, rel grad diff at step 2")
    for (n, off), nxt in zip(names, [o for _, o in names[1:]] + [eng0.n_flat]):
        sl = slice(off, nxt)
        w = float((found[0][1][sl] - ref[0][1][sl]).abs().max())
        nflip = int(((found[0][1][sl] - ref[0][1][sl]).abs() > 1e-6).sum())
        g1 = float((found[0][0][sl] - ref[0][0][sl]).norm() / (ref[0][0][sl].norm() + 1e-30))
        g2 = float((found[1][0][sl] - ref[1][0][sl]).norm() / (ref[1][0][sl].norm() + 1e-30))
        if w > 1e-7:
            print(f"  {n:60s} |dW|max {w:.2e} ({nflip} elems > 1e-6)  g1 {g1:.1e}  g2 {g2:.1e}")
