"""Drop-in for the reference's core/TRANSFORMER/model.py: same import path, same class name, same
constructor / forward / generate_caption_vector / beam_search signatures (model.py:8-209)."""
import os
import sys

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)
import icap_loader  # noqa: E402

_pkg = icap_loader.load()
Transformer = _pkg.Transformer
GraphedTrainStep = _pkg.GraphedTrainStep
PrefetchLoader = _pkg.PrefetchLoader

__all__ = ["Transformer", "GraphedTrainStep", "PrefetchLoader"]
