"""icap_beam_select on decode-like inputs (512 images x 5 beams x 10000 bf16 logits): the two-pass kernel
(ICAP_BEAM_SELECT_NO_STAGE=1) against the shared-memory-staged one, for a few score distributions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402

N = icap_loader.load()._native
dev = torch.device("cuda:0")
B, k, V = 512, 5, 10000
g = torch.Generator(device="cuda").manual_seed(0)
cases = {
    "randn*0.3, prev~U(0,1e-4)": (torch.randn(B * k, V, device=dev, generator=g) * 0.3, torch.rand(B, k, device=dev, generator=g) * 1e-4),
    "randn*0.3, prev~U(0,2e-3) (late step)": (torch.randn(B * k, V, device=dev, generator=g) * 0.3, torch.rand(B, k, device=dev, generator=g) * 2e-3),
    "randn*3, prev~U(0,1)": (torch.randn(B * k, V, device=dev, generator=g) * 3, torch.rand(B, k, device=dev, generator=g)),
    "randn*0.01 (flat)": (torch.randn(B * k, V, device=dev, generator=g) * 0.01, torch.rand(B, k, device=dev, generator=g) * 1e-3),
}
os_ = torch.empty(B, k, device=dev)
op = torch.empty(B, k, dtype=torch.int32, device=dev)
ot = torch.empty(B, k, dtype=torch.int32, device=dev)
gap = torch.empty(B, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, (lg, prev) in cases.items():
    lg = lg.bfloat16().contiguous()
    row = f"{name:42s}"
    for old in ("1", None):
        if old:
            os.environ["ICAP_BEAM_SELECT_NO_STAGE"] = old
        else:
            os.environ.pop("ICAP_BEAM_SELECT_NO_STAGE", None)
        N.call("icap_reload_env")
        ts = []
        for it in range(6):
            flush.zero_() if it % 2 else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.call("icap_beam_select", N.BF16, B, k, V, lg.data_ptr(), V, prev.data_ptr(), k, os_.data_ptr(), op.data_ptr(),
                   ot.data_ptr(), gap.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        row += f"  {'two-pass' if old else 'staged'}: warm {min(ts[0::2][1:]):7.1f} us, cold {min(ts[1::2]):7.1f} us"
    print(row)
