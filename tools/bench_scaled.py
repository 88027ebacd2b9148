"""BASELINE configs[4] (scaled variant) on ONE GPU: model C (d1024 / 16 heads / FFN 4096 / 6+6 blocks, 100 regions x 2048,
vocab 30k, T = 21): training step (batch per GPU 256 unless --batch) and KV-cached beam-5 decode (512 images).
Prints one JSON line; not part of the bench.py contract (its headline is configs[1])."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icap_loader  # noqa: E402
from oracle import caption_oracle as O  # noqa: E402  (synthetic inputs only)

GFLOP_TRAIN, GFLOP_BEAM5 = 69.75, 42.78        # SURVEY.md 8(d), model C


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--decode-batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    pkg = icap_loader.load()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:                                   # data parallel: --batch samples per GPU (weak scaling)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    kw = dict(num_vocab=30000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="scaled",
              dropout=0.2, encode_input_size=1024, encode_q_k_dim=1024, encode_v_dim=1024, encode_hidden_size=4096,
              encode_num_blocks=6, encode_num_heads=16, dim_word_embedding=1024, decode_input_size=1024,
              decode_q_k_dim=1024, decode_v_dim=1024, decode_hidden_size=4096, decode_num_blocks=6, decode_num_heads=16)
    torch.manual_seed(0)
    model = pkg.Transformer(device=dev, **kw).to(dev).train()
    nparam = sum(q.numel() for q in model.parameters())
    f, p, c = O.synthetic_batch(args.batch, 100, 2048, 84, 22, 30000, seed=1 + rank)
    f, p, c = f.to(dev), p.to(dev), c.to(dev)
    dp = pkg.DataParallel(model, dist) if world > 1 else None
    gs = pkg.GraphedTrainStep(model, args.batch, 100, 22, lr=5e-4, dp=dp)
    gs.load(f, p, c)
    gs.capture()
    for _ in range(3):
        gs.step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = gs.step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    out = {"model": "C (d1024/16h/ffn4096/6+6, R=100, V=30k)", "params": nparam, "n_gpus": world,
           "train_batch_per_gpu": args.batch, "train_ms_per_step": ms,
           "train_samples_per_s": world * args.batch / ms * 1e3, "final_loss": float(loss),
           "train_tflops_per_gpu": args.batch / ms * 1e3 * GFLOP_TRAIN / 1e3}
    if world == 1:
        del gs
    model.eval()
    fd, pd, _ = O.synthetic_batch(args.decode_batch, 100, 2048, 84, 22, 30000, seed=2 + rank)
    fd, pd = fd.to(dev), pd.to(dev)
    gd = pkg.GraphedDecode(model, args.decode_batch, 100, 5)       # partitioned by image: every rank its own batch
    for _ in range(2):
        gd.run(fd, pd)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0.record()
    for _ in range(3):
        gd.run(fd, pd)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    msd = float(t)
    out.update({"beam5_batch_per_gpu": args.decode_batch, "beam5_ms": msd,
                "beam5_captions_per_s": world * args.decode_batch / msd * 1e3,
                "beam5_tflops_per_gpu": args.decode_batch / msd * 1e3 * GFLOP_BEAM5 / 1e3})
    if world > 1:
        if rank == 0:
            print(json.dumps(out), flush=True)
        torch.cuda.synchronize()
        os._exit(0)               # no NCCL teardown with captured graphs alive (see bench.py)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
