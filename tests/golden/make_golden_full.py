"""Golden vectors of the UNMODIFIED reference at the BENCHMARKED widths (model A = constructor defaults, BASELINE
configs[1-3]; model B = core/config.py defaults, configs[0]).  The 223 MB / 30 MB state_dicts are not stored: weights and
inputs are regenerated from seeds with the torch CPU generator (oracle.init_state_dict / synthetic_batch; serial, so
independent of the thread count) and pinned by a SHA-256 in the file; what IS stored are the reference's outputs on
them -- loss, a logits slice + whole-tensor sums, per-parameter gradient norms / sums / leading elements, greedy ids and top-2 gaps of every decision, beam-3 / beam-5 ids (first two
images) with the candidate gaps of every beam step.  Seeds were picked so that no greedy decision is a near-tie
(smallest top-2 gap: model A 0.13, model B 3.7e-4).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_full.py        # build container only (/root/reference)

tests/test_oracle.py replays them through the oracle (CPU), tests/test_model_gpu.py through the CUDA path (fp32 mode).
"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

FULL_CASES = {
    # name: (ctor kwargs, batch, regions, weight seed, input seed)
    "modelA": (dict(num_vocab=10000, max_length=22, encode_dim_positions=84, encode_dim_features=2048, output_name="x",
                    dropout=0.0), 4, 36, 11, 4242),
    "modelB": (dict(num_vocab=10000, max_length=51, encode_dim_positions=84, encode_dim_features=2048, output_name="x",
                    dropout=0.0, encode_mask=True, split_image_objects=True, encode_input_size=256, encode_q_k_dim=256,
                    encode_v_dim=256, encode_hidden_size=256, encode_num_blocks=2, encode_num_heads=32,
                    dim_word_embedding=256, decode_input_size=256, decode_q_k_dim=256, decode_v_dim=256,
                    decode_hidden_size=256, decode_num_blocks=5, decode_num_heads=32), 8, 36, 12, 4342),
}


def digest(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def regenerate(name):
    """(cfg, state_dict, features, positions, captions, sha256 of weights + inputs) of a case, from its seeds."""
    from oracle import caption_oracle as O
    kw, B, R, wseed, iseed = FULL_CASES[name]
    cfg = O.OracleConfig(**kw)
    sd = O.init_state_dict(cfg, seed=wseed)
    f, p, c = O.synthetic_batch(B, R, kw["encode_dim_features"], kw["encode_dim_positions"], kw["max_length"],
                                kw["num_vocab"], seed=iseed)
    return cfg, sd, f, p, c, digest(list(sd.values()) + [f, p, c])


def main():
    import make_golden
    from oracle import caption_oracle as O
    Transformer, _ = make_golden.import_reference()
    out = {"torch_version": torch.__version__}
    for name, (kw, B, R, wseed, iseed) in FULL_CASES.items():
        cfg, sd, f, p, c, sha = regenerate(name)
        model = Transformer(device=torch.device("cpu"), **kw)
        model.load_state_dict(sd)
        model.eval()
        grabbed = {}
        h = model.classifer.register_forward_hook(lambda m, i, o: grabbed.__setitem__("logits", o.detach().clone()))
        with torch.no_grad():
            loss = model(f, p, c)["loss"]
        h.remove()
        lg = grabbed["logits"].reshape(B, kw["max_length"] - 1, kw["num_vocab"])
        case = {"ctor": kw, "batch": B, "regions": R, "weight_seed": wseed, "input_seed": iseed, "sha256": sha,
                "loss": loss.detach().clone(), "logits_slice": lg[:, :, :64].clone(),
                "logits_sum": lg.double().sum(), "logits_abs_sum": lg.double().abs().sum(),
                "logits_row_max": lg.max(dim=-1).values.clone(), "logits_row_argmax": lg.argmax(dim=-1).clone()}
        # every gradient of the teacher-forced loss, as (Frobenius norm, sum, first 32 elements) per parameter
        model.zero_grad()
        model(f, p, c)["loss"].backward()
        case["grad_stats"] = {k: (float(q.grad.double().norm()), float(q.grad.double().sum()),
                                  q.grad.reshape(-1)[:32].detach().clone())
                              for k, q in model.named_parameters() if q.grad is not None}
        model.zero_grad()
        # greedy: ids + the top-2 gap of every decision, from the reference's own per-step logits (classifer hook)
        steps = []
        h = model.classifer.register_forward_hook(lambda m, i, o: steps.append(o.detach().clone()))      # [B, V] per decision
        with torch.no_grad():
            ids, _ = model.generate_caption_vector(f, p)
        h.remove()
        top2 = torch.stack([s.topk(2, dim=-1).values for s in steps], dim=1)           # [B, T, 2]
        case["greedy_ids"] = ids.clone()
        case["greedy_gaps"] = (top2[..., 0] - top2[..., 1]).clone()
        with torch.no_grad():
            for k in ((3, 5) if name == "modelA" else (3,)):
                case[f"beam{k}_ids"] = model.beam_search(f[:2], p[:2], beam_size=k).clone()
                # k-th vs (k+1)-th candidate score gap of every beam step (near-tie reporting): the reference does not
                # expose it, so it comes from the oracle AFTER its ids were checked against the reference's
                o_ids, o_trace = O.beam_search(sd, cfg, f[:2], p[:2], beam_size=k, return_trace=True)
                assert torch.equal(o_ids, case[f"beam{k}_ids"]), (name, k)
                case[f"beam{k}_gaps"] = o_trace.clone()
        out[name] = case
        print(name, "loss", float(loss), "min greedy gap", float(case["greedy_gaps"].min()), "sha", sha[:16])
    path = os.path.join(HERE, "full_width.pt")
    torch.save(out, path)
    print("bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
