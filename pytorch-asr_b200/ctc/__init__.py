"""Python surface of the engine, laid out like the reference's wrapper package
around its native module (asr/kaldi/latgen/__init__.py:1-8)."""
from ._ctc import CTCLoss, ctc_loss, ctc_loss_parts, load_native  # noqa: F401
