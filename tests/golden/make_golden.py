"""Generate golden vectors by running the UNMODIFIED reference (read from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these files ARE the parity pin:
each case stores the reference's random-init state_dict, the seeded synthetic inputs, and the
reference's own outputs (loss, logits, every gradient, two Adam steps, greedy ids + attention
rows, beam ids for k=2,3, PolicyNetwork log-domain beam ids).  Cases are tiny so the files stay
small; tests/test_oracle.py replays them through oracle/caption_oracle.py.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def import_reference():
    """Import shim of SURVEY.md Appendix A (stubs for un-vendored pycocoevalcap + hickle)."""
    def _stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    for pkg in ["core.metrics", "core.metrics.cider", "core.metrics.ciderD", "core.metrics.bleu"]:
        _stub(pkg).__path__ = []
    _stub("core.metrics.cider.cider", Cider=_Dummy)
    _stub("core.metrics.ciderD.ciderD", CiderD=_Dummy)
    _stub("core.metrics.bleu.bleu", Bleu=_Dummy)
    _stub("hickle")
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from core.TRANSFORMER.model import Transformer
    from core.TRANSFORMER.model_RL import PolicyNetwork
    return Transformer, PolicyNetwork


CASES = {
    # name: (ctor kwargs, batch, regions)
    "tiny_default": (dict(num_vocab=397, max_length=9, encode_dim_positions=12, encode_dim_features=48,
                          output_name="x", dropout=0.0,
                          encode_input_size=32, encode_q_k_dim=32, encode_v_dim=32, encode_hidden_size=64,
                          encode_num_blocks=2, encode_num_heads=4, dim_word_embedding=24,
                          decode_input_size=32, decode_q_k_dim=32, decode_v_dim=32, decode_hidden_size=64,
                          decode_num_blocks=2, decode_num_heads=4), 5, 6),
    "tiny_cfgpy": (dict(num_vocab=397, max_length=8, encode_dim_positions=12, encode_dim_features=48,
                        output_name="x", dropout=0.0, encode_mask=True, split_image_objects=True,
                        encode_input_size=32, encode_q_k_dim=32, encode_v_dim=32, encode_hidden_size=32,
                        encode_num_blocks=1, encode_num_heads=8, dim_word_embedding=32,
                        decode_input_size=32, decode_q_k_dim=32, decode_v_dim=32, decode_hidden_size=32,
                        decode_num_blocks=2, decode_num_heads=8), 4, 5),
    "tiny_variants": (dict(num_vocab=397, max_length=8, encode_dim_positions=12, encode_dim_features=48,
                           output_name="FocalLoss_x", dropout=0.0, split_position=True,
                           move_first_image_feature=True,
                           encode_input_size=32, encode_q_k_dim=64, encode_v_dim=16, encode_hidden_size=64,
                           encode_num_blocks=1, encode_num_heads=4, dim_word_embedding=32,
                           decode_input_size=32, decode_q_k_dim=64, decode_v_dim=16, decode_hidden_size=64,
                           decode_num_blocks=1, decode_num_heads=4), 4, 5),
}


def main():
    Transformer, PolicyNetwork = import_reference()
    sys.path.insert(0, ROOT)
    from oracle.caption_oracle import synthetic_batch  # input generator only

    for name, (kw, B, R) in CASES.items():
        torch.manual_seed(0)
        model = Transformer(device=torch.device("cpu"), **kw)
        model.eval()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        feats, pos, cap = synthetic_batch(B, R, kw["encode_dim_features"], kw["encode_dim_positions"],
                                          kw["max_length"], kw["num_vocab"], seed=1234)
        feats2, pos2, cap2 = synthetic_batch(B, R, kw["encode_dim_features"], kw["encode_dim_positions"],
                                             kw["max_length"], kw["num_vocab"], seed=99)
        out = {"ctor": kw, "state_dict": sd0, "features": feats, "positions": pos, "captions": cap,
               "features2": feats2, "positions2": pos2, "captions2": cap2,
               "state_dict_keys": list(sd0.keys())}

        # logits via a forward hook on classifer (model.py:93)
        grabbed = {}
        h = model.classifer.register_forward_hook(lambda m, i, o: grabbed.__setitem__("logits", o.detach().clone()))
        loss = model(feats, pos, cap)["loss"]
        h.remove()
        out["loss"] = loss.detach().clone()
        out["logits"] = grabbed["logits"]
        model.zero_grad()
        loss.backward()
        out["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
                        for k, p in model.named_parameters()}

        # decode (before any optimizer step)
        ids, att = model.generate_caption_vector(feats, pos)
        out["greedy_ids"] = ids.clone()
        out["greedy_attention"] = torch.tensor(np.stack(att, 0))
        for k in (2, 3):
            out[f"beam{k}_ids"] = model.beam_search(feats, pos, beam_size=k).clone()

        # PolicyNetwork (same submodule names => same state_dict) : logits + log-domain beam
        pkw = {k: v for k, v in kw.items() if k != "output_name"}
        pol = PolicyNetwork(device=torch.device("cpu"), **pkw)
        pol.load_state_dict(sd0)
        pol.eval()
        with torch.no_grad():
            out["policy_logits"] = pol(feats, pos, cap).clone()
        out["policy_beam3_ids"] = pol.beam_search(feats, pos, beam_size=3).clone()

        # two Adam steps (models.py:111-126), eval-mode arithmetic (dropout = 0 in these cases)
        opt = torch.optim.Adam((p for p in model.parameters() if p.requires_grad), lr=5e-4)
        losses = []
        for f_, p_, c_ in ((feats, pos, cap), (feats2, pos2, cap2)):
            opt.zero_grad()
            l_ = model(f_, p_, c_)["loss"]
            l_.backward()
            opt.step()
            losses.append(float(l_))
        out["adam_losses"] = torch.tensor(losses, dtype=torch.float64)
        out["state_dict_after_2_steps"] = {k: v.detach().clone() for k, v in model.state_dict().items()}
        out["torch_version"] = torch.__version__
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(out, path)
        print(name, "loss", float(out["loss"]), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
