"""Data feed on a B200: a batch named by image numbers into the device-resident RegionCache must be
bit-identical to the same images passed as fp32 tensors the way the reference's DataLoader delivers them
(core/dataset.py:12-18, core/models.py:115-120) -- loss, gradients, the fused training step and decoded ids --
and against the CPU oracle; PrefetchLoader must deliver every batch unchanged and in order."""
import numpy as np
import pytest
import torch

import icap_loader
from oracle import caption_oracle as O

pytestmark = pytest.mark.gpu
pkg = icap_loader.load()
DEV = torch.device("cuda:0")


def small_cfg(**over):
    kw = dict(num_vocab=500, max_length=22, encode_dim_positions=84, encode_dim_features=256, output_name="t",
              encode_num_blocks=1, decode_num_blocks=2, dropout=0.2)
    kw.update(over)
    return kw


def make(kw, precision, seed=0):
    sd = O.init_state_dict(O.OracleConfig(**kw), seed=seed)
    m = pkg.Transformer(device=DEV, **kw)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    m.set_precision(precision)
    return m, sd


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("variant", [{}, {"encode_mask": True}, {"split_image_objects": True}])
def test_cached_batch_is_bit_identical_to_tensor_batch(precision, variant):
    kw = small_cfg(**variant)
    m, sd = make(kw, precision)
    n_img, R = 23, 13                                          # pool of images; ragged region counts inside
    F, P, _ = O.synthetic_batch(n_img, R, 256, 84, 22, 500, seed=3)
    cache = pkg.RegionCache(m, F.numpy(), P.numpy(), chunk_images=5)      # 5 chunks, the last one partial
    assert cache.xcat.dtype == (torch.float32 if precision == "fp32" else torch.bfloat16)
    idx = torch.tensor([4, 4, 0, 22, 9, 17, 4, 1, 9])          # repeats: five captions share an image in COCO
    _, _, C = O.synthetic_batch(len(idx), R, 256, 84, 22, 500, seed=4)
    lg_t = m.logits(F[idx], P[idx], C)
    lg_c = m.logits(cache.batch(idx), None, C)
    assert torch.equal(lg_t, lg_c)
    if precision == "fp32":                                    # and the oracle agrees (north_star: 1e-4 relative)
        ref = O.logits_forward(sd, O.OracleConfig(**kw), F[idx], P[idx], C)
        assert float((lg_c.cpu() - ref).abs().max() / ref.abs().max()) < 1e-4
    loss_t = m(F[idx], P[idx], C)["loss"]
    loss_t.backward()
    g_t = {n: q.grad.clone() for n, q in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    loss_c = m(cache.batch(idx.int()), None, C)["loss"]        # int32 indices too
    loss_c.backward()
    assert abs(float(loss_t.detach()) - float(loss_c.detach())) <= 1e-6 * abs(float(loss_t.detach()))
    for n, q in m.named_parameters():      # same inputs, same kernels: only the order of fp32 atomic / split-K adds varies
        assert float((q.grad - g_t[n]).norm()) <= 1e-5 * float(g_t[n].norm()) + 1e-12, n
    cache.check()


def test_cached_decode_ids_and_attention_identical():
    kw = small_cfg()
    m, _ = make(kw, "fp32")
    F, P, _ = O.synthetic_batch(40, 13, 256, 84, 22, 500, seed=5)
    cache = pkg.RegionCache(m, F, P)                           # CPU tensors accepted as well
    idx = torch.arange(39, 7, -2)
    ids_t, att_t = m.generate_caption_vector(F[idx], P[idx])
    ids_c, att_c = m.generate_caption_vector(cache.batch(idx), None)
    assert torch.equal(ids_t, ids_c)
    # head-mean attention is accumulated with fp32 atomics over the heads: equal up to the order of 8 additions
    assert all(np.allclose(a, b, rtol=0, atol=1e-6) for a, b in zip(att_t, att_c))
    for k in (3, 5):
        assert torch.equal(m.beam_search(F[idx], P[idx], beam_size=k), m.beam_search(cache.batch(idx), None, beam_size=k))
    # a second batch through the SAME captured graph (static index buffer refreshed, not baked in)
    idx2 = torch.arange(0, 16)
    assert torch.equal(m.beam_search(F[idx2], P[idx2], beam_size=3), m.beam_search(cache.batch(idx2), None, beam_size=3))


def test_graphed_train_step_from_cache_equals_tensor_step():
    kw = small_cfg(dropout=0.0)
    F, P, _ = O.synthetic_batch(64, 13, 256, 84, 22, 500, seed=6)
    g = torch.Generator().manual_seed(1)
    batches = []
    for s in range(3):
        idx = torch.randint(0, 64, (32,), generator=g)
        _, _, C = O.synthetic_batch(32, 13, 256, 84, 22, 500, seed=10 + s)
        batches.append((idx, C))
    a, _ = make(kw, "bf16")
    b, _ = make(kw, "bf16")
    cache = pkg.RegionCache(b, F, P, chunk_images=17)
    ga = pkg.GraphedTrainStep(a, 32, 13, 22, lr=5e-4)
    gb = pkg.GraphedTrainStep(b, 32, 13, 22, lr=5e-4, cache=cache)
    # dropout masks are a pure function of (step counter, call site, element), so both models draw the same masks
    for idx, C in batches:
        ga.load(F[idx], P[idx], C)
        gb.load(idx, None, C)
        la, lb = ga.step(), gb.step()
        assert abs(float(la) - float(lb)) <= 1e-4 * abs(float(la))     # a stale index buffer would show up as O(1e-2)
    # Adam's first steps move every weight by ~lr * sign(g): a gradient element at rounding-noise level may flip, so
    # compare whole matrices in norm rather than element by element
    for (n1, q1), (n2, q2) in zip(a.state_dict().items(), b.state_dict().items()):
        if q1.dim() >= 2 and q1.is_floating_point():
            assert float((q1 - q2).norm()) <= 1e-3 * float(q1.norm()), n1
    assert gb.launches_per_step <= ga.launches_per_step      # one gather replaces region_valid + two packing copies
    cache.check()


def test_out_of_range_index_is_reported():
    kw = small_cfg()
    m, _ = make(kw, "bf16")
    F, P, C = O.synthetic_batch(8, 13, 256, 84, 22, 500, seed=7)
    cache = pkg.RegionCache(m, F, P)
    with pytest.raises(IndexError):
        cache.batch(torch.tensor([0, 8]))                      # host-side indices are checked up front
    bad = torch.tensor([0, 8, 3, -1, 2, 2, 2, 2], device=DEV)  # device-side indices: flagged by the kernel
    m.logits(cache.batch(bad), None, C)
    with pytest.raises(IndexError):
        cache.check()
    cache.check()                                              # flag cleared


def test_gather_regions_kernel_direct():
    """C ABI: rows, validity bytes and the row scale for fp32 and bf16 packed widths, int32 and int64 indices."""
    N = pkg._native
    s = torch.cuda.current_stream().cuda_stream
    for dt, code in ((torch.float32, N.F32), (torch.bfloat16, N.BF16)):
        for Kc, R, n, B in ((64, 1, 3, 5), (2176, 37, 11, 19), (192, 100, 4, 7)):
            cache = torch.randn(n, R, Kc, device=DEV).to(dt)
            valid = (torch.rand(n, R, device=DEV) > 0.3).to(torch.uint8)
            for it in (torch.int32, torch.int64):
                idx = torch.randint(0, n, (B,), device=DEV).to(it)
                x = torch.empty(B * R, Kc, dtype=dt, device=DEV)
                kv = torch.empty(B * R, dtype=torch.uint8, device=DEV)
                rs = torch.empty(B * R, dtype=torch.float32, device=DEV)
                err = torch.zeros(1, dtype=torch.int32, device=DEV)
                N.call("icap_gather_regions", code, cache.data_ptr(), valid.data_ptr(), n, idx.data_ptr(),
                       int(it == torch.int64), B, R, Kc, x.data_ptr(), kv.data_ptr(), rs.data_ptr(), err.data_ptr(), s)
                assert torch.equal(x.view(B, R, Kc), cache[idx.long()])
                assert torch.equal(kv.view(B, R), valid[idx.long()])
                assert torch.equal(rs.view(B, R), valid[idx.long()].float())
                assert int(err) == 0
    with pytest.raises(N.IcapError):                           # 24-byte rows cannot be moved in 16-byte lanes
        N.call("icap_gather_regions", N.BF16, cache.data_ptr(), valid.data_ptr(), 1, idx.data_ptr(), 1, 1, 1, 12,
               x.data_ptr(), kv.data_ptr(), rs.data_ptr(), None, s)


def test_prefetch_loader_delivers_batches_in_order():
    g = torch.Generator().manual_seed(0)
    data = [(torch.randn(4, 7, 16, generator=g), torch.randint(0, 9, (4, 5), generator=g, dtype=torch.int32))
            for _ in range(7)]
    data.append((torch.randn(2, 7, 16, generator=g), torch.randint(0, 9, (2, 5), generator=g, dtype=torch.int32)))  # ragged tail
    for depth in (2, 3):
        got = []
        for f, c in pkg.PrefetchLoader(data, DEV, depth=depth):
            assert f.is_cuda and c.is_cuda
            got.append((f.clone(), c.clone()))                 # consumed on the compute stream before the slot is reused
        assert len(got) == len(data)
        for (f, c), (rf, rc) in zip(got, data):
            assert torch.equal(f.cpu(), rf) and torch.equal(c.cpu(), rc)
    assert list(pkg.PrefetchLoader([], DEV)) == []
