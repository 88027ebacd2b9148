"""Data parallel + device-resident region cache on N GPUs (torchrun): every rank feeds its own image numbers from the
same cache; after 3 steps all ranks hold identical weights and the losses equal those of the tensor-fed DP step.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_cache_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402  (synthetic inputs only)

pkg = icap_loader.load()
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
kw = dict(num_vocab=2000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="t", dropout=0.1)
F, P, _ = O.synthetic_batch(96, 36, 2048, 84, 22, 2000, seed=7)
losses = {}
finals = {}
for mode in ("tensor", "cache"):
    torch.manual_seed(0)
    m = pkg.Transformer(device=dev, **kw).to(dev).train()
    dp = pkg.DataParallel(m, dist)
    cache = pkg.RegionCache(m, F, P) if mode == "cache" else None
    gs = pkg.GraphedTrainStep(m, 64, 36, 22, lr=5e-4, dp=dp, cache=cache)
    g = torch.Generator().manual_seed(100 + rank)
    out = []
    for s in range(3):
        idx = torch.randint(0, 96, (64,), generator=g)
        _, _, C = O.synthetic_batch(64, 36, 8, 84, 22, 2000, seed=1000 * rank + s)
        if cache is None:
            gs.load(F[idx], P[idx], C)
        else:
            gs.load(idx, None, C)
        out.append(float(gs.step()))
    losses[mode] = out
    w = m._engine().p32.double()
    finals[mode] = (float(w.sum()), float(w.abs().sum()))
    ws = torch.tensor([finals[mode][0]], device=dev, dtype=torch.float64)
    lo, hi = ws.clone(), ws.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(hi - lo) == 0.0, f"ranks diverged in {mode} mode: {float(lo)} vs {float(hi)}"
    if cache is not None:
        cache.check()
    del gs, dp, m
for a, b in zip(losses["tensor"], losses["cache"]):
    assert abs(a - b) <= 2e-4 * abs(a), (losses, rank)
if rank == 0:
    print("OK", world, "ranks; losses tensor-fed", losses["tensor"], "cache-fed", losses["cache"], flush=True)
torch.cuda.synchronize()
os._exit(0)          # no NCCL teardown with captured graphs alive (see bench.py)
