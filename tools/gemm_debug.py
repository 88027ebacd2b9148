"""Debug helper: error map of icap_gemm (bf16) cases; prints which 32x32 blocks of C are wrong."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

pkg = icap_loader.load()
N = pkg._native
F32, BF16 = N.F32, N.BF16
dev = torch.device("cuda:0")


def case(ak, bk, M, Nn, K, cdt, bias, epi, acc, split, tag=""):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((M, K) if ak else (K, M), device=dev, generator=g).bfloat16()
    B = torch.randn((Nn, K) if bk else (K, Nn), device=dev, generator=g).bfloat16()
    ref = (A if ak else A.t()).double() @ (B.t() if bk else B).double()
    tdt = torch.float32 if cdt == F32 else torch.bfloat16
    bias_t = torch.randn(Nn, device=dev, generator=g) if bias else None
    if bias:
        ref = ref + bias_t.double()
    aux = None
    if epi == 1:
        ref = ref.clamp_min(0)
    if epi == 2:
        aux = torch.randn(M, Nn, device=dev, generator=g).to(tdt)
        ref = ref * (aux.double() > 0)
    C = torch.randn(M, Nn, device=dev, generator=g).to(tdt) if acc else torch.full((M, Nn), float("nan"), device=dev, dtype=tdt)
    if acc:
        ref = ref + C.double()
    N.call("icap_gemm", BF16, ak, bk, M, Nn, K, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], C.data_ptr(), Nn, cdt,
           bias_t.data_ptr() if bias else None, epi, aux.data_ptr() if aux is not None else None, Nn, acc, split,
           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = (C.double() - ref).abs()
    err = torch.nan_to_num(err, nan=1e9)
    tol = (1e-4 if cdt == F32 else 2e-2) * float(ref.abs().max())
    bad = err > tol
    msg = f"{tag} ak{ak} bk{bk} {M}x{Nn}x{K} c={'f32' if cdt == F32 else 'bf16'} bias{bias} epi{epi} acc{acc} split{split} " \
          f"BN={os.environ.get('ICAP_GEMM_BN', 'auto')}: max err {float(err.max()):.3g} (tol {tol:.3g}) bad {int(bad.sum())}/{bad.numel()}"
    print(msg)
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print("   bad rows: min %d max %d count %d ; bad cols: min %d max %d count %d" %
              (rows.min(), rows.max(), rows.numel(), cols.min(), cols.max(), cols.numel()))
        # 128-row x 32-col block map (first 16 row tiles)
        Mb, Nb = (M + 127) // 128, (Nn + 31) // 32
        for mb in range(min(Mb, 40)):
            line = ""
            for nb in range(Nb):
                blk = bad[mb * 128:(mb + 1) * 128, nb * 32:(nb + 1) * 32]
                fr = float(blk.float().mean()) if blk.numel() else 0.0
                line += "." if fr == 0 else ("#" if fr > 0.9 else "+")
            if "#" in line or "+" in line:
                print(f"   mtile {mb:3d}: {line}")


if __name__ == "__main__":
    for bn in ("128", "256"):
        os.environ["ICAP_GEMM_BN"] = bn
        case(1, 0, 4000, 1000, 200, F32, 0, 0, 1, 1, "fail?")
        case(1, 0, 4000, 1000, 200, F32, 0, 0, 0, 1, "store")
        case(1, 1, 4000, 1000, 200, F32, 0, 0, 1, 1, "kk-acc")
        case(1, 0, 4000, 1024, 256, F32, 0, 0, 1, 1, "aligned")
        case(1, 0, 1000, 1000, 200, F32, 0, 0, 1, 1, "fewtiles")
        case(1, 0, 4000, 1000, 200, BF16, 0, 0, 1, 1, "bf16acc")
