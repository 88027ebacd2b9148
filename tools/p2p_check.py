"""Peer-memory all-reduce (csrc/p2p.cu, transformer.PeerReduce) on N GPUs under torchrun: result == NCCL all-reduce
(bit-identical across ranks), and time per 223 MB buffer alone on the stream."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

pkg = icap_loader.load()
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)


class FakeEng:
    def __init__(self, n):
        self.n_flat, self.dev = n, dev


n = 55_707_408 // 8 * 8
pr = pkg.transformer.PeerReduce(dist, FakeEng(n))
g = torch.Generator(device=dev).manual_seed(100 + rank)
x = torch.randn(n + 8, device=dev, generator=g)
ref = x.clone()
dist.all_reduce(ref)
st = torch.cuda.current_stream().cuda_stream
if rank == 0:
    print(f"multicast pointer {pr.mc:#x}, NVLS path {'on' if pr.use_nvls else 'off'}", flush=True)
for nb in (1, 5):
    pr.g32.copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    bounds = [(n + 8) * i // nb // 8 * 8 for i in range(nb)] + [n + 8]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        pr.all_reduce(lo, hi, st)
    pr.barrier(st)
    torch.cuda.synchronize()
    pr.check()
    err = float((pr.g32 - ref).abs().max() / ref.abs().max())
    gathered = [torch.empty(1, device=dev) for _ in range(world)]
    dist.all_gather(gathered, pr.g32.double().sum().float().reshape(1))
    same = all(float(t) == float(gathered[0]) for t in gathered)
    if rank == 0:
        print(f"{nb} bucket(s): max rel err vs NCCL {err:.2e}, checksums identical on all ranks: {same}", flush=True)
# timing
def two_shot():
    keep, pr.use_nvls = pr.use_nvls, False
    pr.all_reduce(0, n + 8, st)
    pr.barrier(st)
    pr.use_nvls = keep


cases = [("peer kernels, two-shot loads / stores", two_shot), ("nccl", lambda: dist.all_reduce(x))]
if pr.use_nvls:
    cases.insert(0, ("peer kernel, NVSwitch multimem", lambda: (pr.all_reduce(0, n + 8, st), pr.barrier(st))))
for name, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        ms = e0.elapsed_time(e1) / 10
        print(f"{name}: {ms:.3f} ms per all-reduce of {4 * (n + 8) / 1e6:.0f} MB on {world} GPUs "
              f"({2 * (world - 1) / world * 4 * (n + 8) / ms / 1e6:.0f} GB/s bus bandwidth)", flush=True)
pr.check()
torch.cuda.synchronize()
os._exit(0)
