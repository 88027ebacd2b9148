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
