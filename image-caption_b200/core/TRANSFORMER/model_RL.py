"""Drop-in for the reference's core/TRANSFORMER/model_RL.py: `PolicyNetwork` with the same constructor, `forward`
(teacher-forced logits), `sample`, `generate_caption_vector` and `beam_search` (log-softmax scores) signatures
(model_RL.py:10-213), executing on libicap.so."""
import os
import sys

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)
import icap_loader  # noqa: E402

PolicyNetwork = icap_loader.load().PolicyNetwork

__all__ = ["PolicyNetwork"]
