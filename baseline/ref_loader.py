"""Import the UNMODIFIED reference classes of the hot path (core/TRANSFORMER/model.py::Transformer,
model_RL.py::PolicyNetwork) for the reference arm of bench.py and the CPU baselines.

The reference is a plain Python tree (no setup.py / pyproject, so `pip install /root/reference` does not apply); its six
files on the path are STAGED, byte for byte, into the git-ignored `baseline/_ref/` by `stage()` (called from
`__graft_entry__.build()` when /root/reference is mounted) so that they travel to the GPU box with the snapshot.
Two of its top-level imports are un-vendored third-party packages that the path never calls (pycocoevalcap, hickle):
they are stubbed exactly as in SURVEY.md Appendix A.  Nothing here is imported by the product package.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
FILES = ["core/__init__.py", "core/utils.py", "core/TRANSFORMER/model.py", "core/TRANSFORMER/modules.py",
         "core/TRANSFORMER/loss.py", "core/TRANSFORMER/model_RL.py"]


def stage() -> bool:
    """Copy the reference's files of the path into baseline/_ref (git-ignored).  Returns True when staged."""
    if not os.path.isdir(REFERENCE):
        return os.path.isdir(STAGED)
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(STAGED, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


def available() -> bool:
    return os.path.exists(os.path.join(STAGED, "core", "TRANSFORMER", "model.py"))


_loaded = None


def load():
    """(Transformer, PolicyNetwork) of the staged reference, or None when it is not staged."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        return None

    def _stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:            # stands in for the pycocoevalcap scorers; never called on the hot path
        def __init__(self, *a, **k):
            pass

    for pkg in ["core.metrics", "core.metrics.cider", "core.metrics.ciderD", "core.metrics.bleu"]:
        _stub(pkg).__path__ = []
    _stub("core.metrics.cider.cider", Cider=_Dummy)
    _stub("core.metrics.ciderD.ciderD", CiderD=_Dummy)
    _stub("core.metrics.bleu.bleu", Bleu=_Dummy)
    if "hickle" not in sys.modules:
        _stub("hickle")
    sys.dont_write_bytecode = True
    # the product package mirrors the reference's tree under another top-level name, so `core` is free
    for k in [k for k in sys.modules if k == "core" or k.startswith("core.") and not k.startswith("core.metrics")]:
        del sys.modules[k]
    sys.path.insert(0, STAGED)
    try:
        from core.TRANSFORMER.model import Transformer
        from core.TRANSFORMER.model_RL import PolicyNetwork
    finally:
        sys.path.remove(STAGED)
    _loaded = (Transformer, PolicyNetwork)
    return _loaded
