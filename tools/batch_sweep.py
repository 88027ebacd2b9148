"""Per-GPU compute time of BASELINE configs[3] (model A, GLOBAL batch 2048): the graphed training step (forward +
backward + Adam, bf16, dropout on) at the per-GPU batches of its 8 / 4 / 2 / 1-GPU splits (256 / 512 / 1024 / 2048),
timed on ONE GPU with inputs resident in HBM.  The data-parallel exchange adds a fixed ~0.3 ms per step on top of
these (profiles/r2_summary.md section 5), so the table is the strong-scaling curve minus that constant.

    python tools/batch_sweep.py [batch ...]        # default: 2048 1024 512 256
    python tools/batch_sweep.py decode [images ...]  # graphed KV-cached beam-5 decode at other batch sizes (default 2048 1024 256)
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402   (synthetic input generator only)

GFLOP_TRAIN = 8.458          # algorithmic matmul GFLOP per sample, model A, fwd + bwd
GFLOP_BEAM5 = 7.274          # the same per image for the KV-cached beam-5 decode (21 steps)
pkg = icap_loader.load()
dev = torch.device("cuda:0")
_pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peaks = json.load(open(_pk)) if os.path.exists(_pk) else {}
peak = peaks.get("bf16_tflops", 1590.0)
kw = dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="bench", dropout=0.2)
decode = len(sys.argv) > 1 and sys.argv[1] == "decode"
batches = [int(a) for a in sys.argv[2 if decode else 1:]] or ([2048, 1024, 256] if decode else [2048, 1024, 512, 256])
steps = int(os.environ.get("SWEEP_STEPS", "20"))
torch.manual_seed(0)
model = pkg.Transformer(device=dev, **kw).to(dev).train()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if decode:
    model.eval()
    print(f"model A KV-cached beam-5 decode on one B200 (whole graph: encoder + cross-K/V + 21 steps), 5 timed replays after 2 "
          f"warm-ups; fractions of the burst bf16 peak {peak:.1f} TFLOP/s", flush=True)
    for B in batches:
        f, p, _ = O.synthetic_batch(B, 36, 2048, 84, 22, 10000, seed=4321)
        f, p = f.to(dev), p.to(dev)
        gd = pkg.GraphedDecode(model, B, 36, 5)
        for _ in range(2):
            gd.run(f, p)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(5):
            gd.run(f, p)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 5
        cps = B / (ms / 1e3)
        tf = cps * GFLOP_BEAM5 / 1e3
        print(json.dumps({"images": B, "decoder_rows": 5 * B, "ms_per_batch": round(ms, 3), "captions_per_s": round(cps, 1),
                          "model_tflops": round(tf, 1), "frac_of_burst_peak": round(tf / peak, 4)}), flush=True)
        del gd, f, p
        torch.cuda.empty_cache()
    sys.exit(0)
print(f"model A train step on one B200, {steps} timed graph replays after 5 warm-ups, 2 rotating input batches; "
      f"fractions of the burst bf16 peak {peak:.1f} TFLOP/s", flush=True)
for B in batches:
    t0 = time.perf_counter()
    pool = []
    for i in range(2):
        f, p, c = O.synthetic_batch(B, 36, 2048, 84, 22, 10000, seed=1234 + i)
        pool.append((f.to(dev), p.to(dev), c.to(dev)))
    gs = pkg.GraphedTrainStep(model, B, 36, 22, lr=5e-4)
    gs.load(*pool[0])
    gs.capture()
    for i in range(5):
        gs.load(*pool[i % 2])
        gs.step()
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(steps):
        gs.load(*pool[i % 2])
        loss = gs.step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    sps = B / (ms / 1e3)
    tf = sps * GFLOP_TRAIN / 1e3
    print(json.dumps({"batch_per_gpu": B, "gpus_for_global_2048": 2048 // B, "ms_per_step": round(ms, 4),
                      "samples_per_s": round(sps, 1), "model_tflops": round(tf, 1), "frac_of_burst_peak": round(tf / peak, 4),
                      "launches_per_step": gs.launches_per_step, "loss": round(float(loss), 4),
                      "wall_s_incl_capture": round(time.perf_counter() - t0, 1)}), flush=True)
    del gs, pool
    torch.cuda.empty_cache()
