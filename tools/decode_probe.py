"""Launch the decode-step kernels once with configs[2] shapes (512 images x 5 beams) -- target for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

pkg = icap_loader.load()
N = pkg._native
BF16, F32 = N.BF16, N.F32
dev = torch.device("cuda:0")
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731

B, k, H, T, R, V, d = 512, 5, 8, 21, 36, 10000, 512
rows = B * k
g = torch.Generator(device="cuda").manual_seed(0)
logits = torch.randn(rows, V, device=dev, generator=g).bfloat16()
prev = torch.rand(B, k, device=dev, generator=g)
osc = torch.empty(B, k, device=dev)
opar = torch.empty(B, k, dtype=torch.int32, device=dev)
otok = torch.empty(B, k, dtype=torch.int32, device=dev)
for _ in range(3):
    N.call("icap_beam_select", BF16, B, k, V, logits.data_ptr(), V, prev.data_ptr(), k, osc.data_ptr(), opar.data_ptr(),
           otok.data_ptr(), None, 0, None, 0, S())
# cross attention G = 5
q = torch.randn(rows, d, device=dev, generator=g).bfloat16()
kv = torch.randn(B * R, 2 * d, device=dev, generator=g).bfloat16()
kvalid = torch.ones(B * R, dtype=torch.uint8, device=dev)
o = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    N.call("icap_mha_decode", BF16, rows, H, R, 64, 64, q.data_ptr(), d, kv.data_ptr(), 2 * d, kv.data_ptr() + 2 * d, 2 * d, R,
           o.data_ptr(), d, None, 0, None, 0, 0, kvalid.data_ptr(), k, None, S())
# self attention over the cache, t = 15
t = 15
cache = torch.randn(rows, T, 2 * d, device=dev, generator=g).bfloat16()
tok = torch.randint(1, V, (rows, T + 1), device=dev, generator=g, dtype=torch.int32)
slot = torch.randint(0, rows, (rows, T + 1), device=dev, generator=g, dtype=torch.int32)
for _ in range(3):
    N.call("icap_mha_decode", BF16, rows, H, t + 1, 64, 64, q.data_ptr(), d, cache.data_ptr(), 2 * d, cache.data_ptr() + 2 * d,
           2 * d, T, o.data_ptr(), d, slot.data_ptr(), T + 1, tok.data_ptr(), T + 1, 0, None, 1, None, S())
# the small decode GEMMs
x = torch.randn(rows, d, device=dev, generator=g).bfloat16()
w = torch.randn(3 * d, d, device=dev, generator=g).bfloat16()
y = torch.empty(rows, 3 * d, device=dev, dtype=torch.bfloat16)
for n in (3 * d, d):
    for _ in range(3):
        N.call("icap_gemm", BF16, 1, 1, rows, n, d, x.data_ptr(), d, w.data_ptr(), d, y.data_ptr(), n, BF16, None, 0, None, 0,
               0, 1, S())
torch.cuda.synchronize()
print("ok")
