"""world_size-2 `gloo` test (CPU) of the data-parallel host logic: every rank deposits the
gradient of the SUM of its token losses plus its non-pad token count in the flat [grads | count] buffer;
one SUM all-reduce + division by the reduced count must equal the gradient of the reference's global
mean loss on the concatenated batch.  Per-rank gradients come from the CPU oracle (the CUDA engine
produces the same quantities on a GPU box; see tests/test_model_gpu.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import icap_loader
    from oracle import caption_oracle as O
    pkg = icap_loader.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(num_vocab=120, max_length=9, encode_dim_positions=12, encode_dim_features=24, output_name="x", dropout=0.0,
              encode_input_size=32, encode_q_k_dim=32, encode_v_dim=32, encode_hidden_size=64, encode_num_blocks=1,
              encode_num_heads=4, dim_word_embedding=32, decode_input_size=32, decode_q_k_dim=32, decode_v_dim=32,
              decode_hidden_size=64, decode_num_blocks=1, decode_num_heads=4)
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=3)
    # ragged shards: rank 0 gets 3 captions, rank 1 gets 5 (different non-pad token counts)
    f, p, c = O.synthetic_batch(8, 5, 24, 12, 9, 120, seed=21)
    lo, hi = (0, 3) if rank == 0 else (3, 8)
    mcfg = pkg.ModelConfig(**kw)
    shapes = pkg.param_layout(mcfg)
    offsets, n_flat = pkg.flat_offsets(shapes)
    loss, grads = O.loss_and_grads(sd, cfg, f[lo:hi], p[lo:hi], c[lo:hi])
    count = float((c[lo:hi, 1:] != 0).sum())
    flat = torch.zeros(n_flat + 8)
    for k, off in offsets.items():
        flat[off:off + grads[k].numel()] = (grads[k] * count).reshape(-1)     # gradient of the SUM of token losses
    flat[n_flat] = count
    total = pkg.DataParallel.allreduce_flat(dist, flat, n_flat)
    flat[:n_flat] /= total
    if rank == 0:
        _, ref = O.loss_and_grads(sd, cfg, f, p, c)
        worst = 0.0
        for k, off in offsets.items():
            got = flat[off:off + ref[k].numel()].view_as(ref[k])
            worst = max(worst, float((got - ref[k]).abs().max() / (ref[k].abs().max() + 1e-12)))
        q.put((float(total), float((c[:, 1:] != 0).sum()), worst))
    dist.barrier()
    dist.destroy_process_group()


def _bucket_worker(rank, world, port, q):
    """Bucketed all-reduce in backward-completion order == one all-reduce of the whole buffer."""
    sys.path.insert(0, ROOT)
    import icap_loader
    pkg = icap_loader.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    n = 10_000 + 8
    flat = torch.randn(n, generator=g)
    whole = flat.clone()
    dist.all_reduce(whole)
    plan = pkg.GradBuckets(n, 1500)
    fired = []
    # closures finish with descending (and a few out-of-order / None) lowest offsets
    for lo in [9500, None, 9000, 9700, 8400, 8399, 6000, 6100, None, 3000, 2999, 100]:
        sl = plan.on_done(lo)
        if sl is not None:
            dist.all_reduce(flat[sl[0]:sl[1]])
            fired.append(sl)
    sl = plan.flush()
    dist.all_reduce(flat[sl[0]:sl[1]])
    fired.append(sl)
    assert plan.flush() is None
    covered = sorted(fired)
    ok_cover = covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    if rank == 0:
        q.put((bool(torch.equal(flat, whole)), ok_cover, len(fired)))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_bucketed_allreduce_equals_whole_buffer():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    same, cover, nb = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert same and cover and nb >= 3


def test_dp_allreduce_reproduces_global_mean_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    total, expect, worst = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert total == expect
    assert worst < 1e-4
