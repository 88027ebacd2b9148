"""Do independent chains of dependent small GEMMs overlap?  One chain of `rows` rows vs n chains of rows/n rows on n
streams (forked/joined inside ONE CUDA graph, as GraphedDecode does with ICAP_DECODE_STREAMS), PDL on and off.
    python tools/chain_concurrency.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

pkg = icap_loader.load()
N = pkg._native
BF16 = N.BF16
dev = torch.device("cuda:0")
d, F, LAYERS = 512, 2048, 6
g = torch.Generator(device="cuda").manual_seed(0)
wqkv = (torch.randn(3 * d, d, device=dev, generator=g) * 0.04).bfloat16()
wo = (torch.randn(d, d, device=dev, generator=g) * 0.04).bfloat16()
w1 = (torch.randn(F, d, device=dev, generator=g) * 0.04).bfloat16()
w2 = (torch.randn(d, F, device=dev, generator=g) * 0.02).bfloat16()
b1 = torch.zeros(F, device=dev)


class Chain:
    def __init__(self, M):
        self.M = M
        self.x = torch.randn(M, d, device=dev).bfloat16()
        self.qkv = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
        self.o = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
        self.h = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
        self.y = torch.empty(M, d, device=dev, dtype=torch.bfloat16)

    def gemm(self, a, lda, w, n, k, c, bias=None, epi=0):
        N.call("icap_gemm", BF16, 1, 1, self.M, n, k, a.data_ptr(), lda, w.data_ptr(), k, c.data_ptr(), n, BF16,
               bias.data_ptr() if bias is not None else None, epi | N.EPI_B_STATIC, None, 0, 0, 1,
               torch.cuda.current_stream().cuda_stream)

    def run(self):
        a = self.x
        for _ in range(LAYERS):
            self.gemm(a, d, wqkv, 3 * d, d, self.qkv)
            self.gemm(self.qkv, 3 * d, wo, d, d, self.o)
            self.gemm(self.o, d, w1, F, d, self.h, bias=b1, epi=1)
            self.gemm(self.h, F, w2, d, F, self.y)
            a = self.y


def timed(nchains, rows, pdl):
    N.call("icap_set_pdl", pdl)
    chains = [Chain(rows // nchains) for _ in range(nchains)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nchains)]

    def go():
        main = torch.cuda.current_stream(dev)
        if nchains == 1:
            chains[0].run()
            return
        for c, s in zip(chains, streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                c.run()
        for s in streams:
            main.wait_stream(s)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        go()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        go()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 20


for small in ("0", "1"):
    os.environ["ICAP_GEMM_SMALL"] = small
    N.call("icap_reload_env")
    for pdl in (1, 0):
        row = [f"{timed(n, 2560, pdl):8.1f}" for n in (1, 2, 4)]
        print(f"ICAP_GEMM_SMALL={small} pdl={pdl}: 24-launch chain(s), 2560 rows total, 1 / 2 / 4 chains: " + " / ".join(row) + " us")

# ---- do the two chains really run at the same time?  entry / exit stamps of every launch of a 2-chain graph
os.environ["ICAP_GEMM_SMALL"] = "1"
N.call("icap_reload_env")
N.call("icap_set_pdl", 1)
chains = [Chain(1280) for _ in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
buf = torch.zeros(48, 16, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
N.call("icap_debug_trace", buf.data_ptr(), 48)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    main = torch.cuda.current_stream(dev)
    for c, s in zip(chains, streams):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            c.run()
    for s in streams:
        main.wait_stream(s)
N.call("icap_debug_trace", None, 0)
gr.replay(); gr.replay()
torch.cuda.synchronize()
t = buf.cpu()
base = int(t[:, 0][t[:, 0] > 0].min())
print("2 chains x 24 launches in one graph: [entry, dep_ok, exit] ns of CTA 0 of every launch")
for i in range(24):
    a, b = t[i], t[24 + i]
    print(f"  launch {i:2d}: chain0 [{int(a[0]) - base:7d} {int(a[2]) - base:7d} {int(a[8]) - base:7d}]   "
          f"chain1 [{int(b[0]) - base:7d} {int(b[2]) - base:7d} {int(b[8]) - base:7d}]")
