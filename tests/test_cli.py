"""Entry points `main.py train | evaluation | demo --beam-size` (reference main.py:25,156,193)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "image-caption_b200")


def _run(args, tmp_path, extra_env=None):
    env = dict(os.environ, ICAP_MAX_LENGTH="10", ICAP_SYNTHETIC_VOCAB="500", ICAP_BATCH_SIZE="8", ICAP_NUM_EPOCH="1")
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(PKG, "main.py"), *args], cwd=tmp_path, env=env,
                          capture_output=True, text=True, timeout=600)


def test_fire_shim_parses_reference_style_flags():
    sys.path.insert(0, PKG)
    try:
        import importlib
        main = importlib.import_module("main")
        seen = {}
        orig = main.evaluation
        main.evaluation = lambda **kw: seen.update(kw)        # _fire resolves commands through the module globals
        try:
            assert main._fire(["evaluation", "--split", "test", "--epoch=7", "--beam-size", "5"]) == 0
            assert seen == {"split": "test", "epoch": 7, "beam_size": 5}
            assert main._fire(["bogus"]) == 2
        finally:
            main.evaluation = orig
    finally:
        sys.path.remove(PKG)
        for m in [k for k in sys.modules if k == "main" or k.startswith("core")]:
            sys.modules.pop(m)


@pytest.mark.gpu
def test_train_evaluation_demo_synthetic(tmp_path):
    r = _run(["train", "--num-images", "16", "--max-iters", "3"], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "loss" in r.stdout
    assert os.path.exists(os.path.join(tmp_path, "output"))
    r = _run(["evaluation", "--split", "test", "--epoch", "1", "--beam-size", "3", "--num-images", "16"], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "16 captions" in r.stdout
    r = _run(["demo", "--epoch", "1", "--beam-size", "5"], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Generated Caption:" in r.stdout
    r = _run(["demo", "--epoch", "1"], tmp_path)
    assert r.returncode == 0 and "Generated Caption:" in r.stdout


@pytest.mark.gpu
def test_train_rl_transformer_synthetic(tmp_path):
    """CAPTION_MODEL='RL_Transformer' (the reference's shipped default, config.py:14): SelfCriticNetwork = PolicyNetwork
    logits + log-softmax/arg-max sampler + self-critical loss with the injectable reward, through the same CLI."""
    env = {"ICAP_CAPTION_MODEL": "RL_Transformer"}
    r = _run(["train", "--num-images", "16", "--max-iters", "3"], tmp_path, env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "loss" in r.stdout
    r = _run(["evaluation", "--split", "test", "--epoch", "1", "--beam-size", "3", "--num-images", "16"], tmp_path, env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "16 captions" in r.stdout


def _write_split(root, split, n_img, R, Dp, L, vocab, seed, caps_per_image=2):
    """data/<MODEL_NAME>/<split>/ in the reference's layout (utils.py:32-64), arrays as .npy (hickle is not installed)."""
    import pickle
    import numpy as np
    from oracle import caption_oracle as O
    F, P, _ = O.synthetic_batch(n_img, R, 2048, Dp, L, vocab, seed=seed)
    _, _, C = O.synthetic_batch(n_img * caps_per_image, R, 8, Dp, L, vocab, seed=seed + 1)
    d = os.path.join(root, split)
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, f"{split}.features.npy"), F.numpy())
    np.save(os.path.join(d, f"{split}.positions.npy"), P.numpy())
    idxs = np.repeat(np.arange(n_img), caps_per_image)
    for name, obj in (("file.names", [f"img{i}.jpg" for i in range(n_img)]), ("captions", C.numpy()),
                      ("image.indices", idxs)):
        with open(os.path.join(d, f"{split}.{name}.pkl"), "wb") as f:
            pickle.dump(obj, f)
    if split == "train":
        w = {"<NULL>": 0, "<START>": 1, "<END>": 2, "<UNK>": 3}
        w.update({f"w{i}": i for i in range(4, vocab)})
        with open(os.path.join(d, "word_index.pkl"), "wb") as f:
            pickle.dump(w, f)
    return F, P, C.numpy(), idxs


def test_load_coco_data_and_datasets(tmp_path, monkeypatch):
    """Host side of the data feed (no GPU): the reference's split layout is read back item for item, and
    IndexedCaptions names the same items by image number only."""
    import numpy as np
    F, P, C, idxs = _write_split(str(tmp_path), "train", 6, 5, 84, 12, 50, seed=3)
    sys.path.insert(0, PKG)
    try:
        for m in [k for k in sys.modules if k.startswith("core")]:
            sys.modules.pop(m)
        from core.utils import load_coco_data
        import core.dataset as D
        data = load_coco_data(data_path=str(tmp_path), split="train")
        assert set(data) == {"features", "positions", "file_names", "captions", "image_idxs", "word_to_idx"}
        ds = D.TrainDataset(data_path=str(tmp_path), split="train")
        assert len(ds) == 12 and ds.len_image == 6 and ds.data_dict is ds.data
        f, p, c, j = ds[7]
        assert j == idxs[7] and np.array_equal(f, F[j].numpy()) and np.array_equal(p, P[j].numpy()) and np.array_equal(c, C[7])
        ix = D.IndexedCaptions(ds)
        assert len(ix) == 12 and ix[7][0] == j and np.array_equal(ix[7][1], C[7])
        uq = D.IndexedCaptions(ds, with_captions=False, unique_images=True)
        assert [uq[i][0] for i in range(len(uq))] == list(range(6))
        with pytest.raises(FileNotFoundError):
            load_coco_data(data_path=str(tmp_path), split="valid")
    finally:
        sys.path.remove(PKG)
        for m in [k for k in sys.modules if k.startswith("core")]:
            sys.modules.pop(m)


@pytest.mark.gpu
def test_region_cache_cli_matches_tensor_feed_on_reference_layout_data(tmp_path):
    """train + evaluation over data/<MODEL_NAME>/{train,valid,test} files: the cached feed (default) and the reference's
    tensor feed (--region-cache False) write identical candidate captions from the same checkpoint."""
    import pickle
    env = {"ICAP_OUTPUT_NAME": "ctor_defaults", "ICAP_MAX_LENGTH": "10"}
    data_root = os.path.join(tmp_path, "data", "maxlen49_36obj_1wordCount")
    for split, n, seed in (("train", 12, 1), ("valid", 5, 2), ("test", 9, 3)):
        _write_split(data_root, split, n, 37, 84, 12, 300, seed=seed)
    r = _run(["train", "--max-iters", "4"], tmp_path, env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "region cache" in r.stdout and "valid" in r.stdout
    assert os.path.exists(os.path.join(tmp_path, "output", "ctor_defaults", "model", "optimizer_1.pt"))
    outs = []
    for flag in ("True", "False"):
        r = _run(["evaluation", "--split", "test", "--epoch", "1", "--beam-size", "3", "--region-cache", flag], tmp_path, env)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "9 captions" in r.stdout
        with open(os.path.join(data_root, "test", "ctor_defaults", "test.candidate.captions.pkl"), "rb") as f:
            outs.append(pickle.load(f))
    assert outs[0] == outs[1] and all(isinstance(c, str) and c for c in outs[0])
    r = _run(["train", "--max-iters", "2", "--region-cache", "False"], tmp_path, env)
    assert r.returncode == 0, r.stderr[-2000:]
