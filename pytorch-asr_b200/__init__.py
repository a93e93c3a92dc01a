"""pytorch-asr_b200 -- B200-native CTC loss engine for jinserk/pytorch-asr.

One component, one path: the loss object of `asr/models/trainer.py:152-154`
(`nn.CTCLoss(blank=0, reduction='mean')`) and its call / backward at
`trainer.py:422,438` (`:508,517`), rebuilt as hand-written sm_100a CUDA behind
the same Python surface:

    from pytorch_asr_b200 import CTCLoss
    trainer = NonSplitTrainer(model, loss=CTCLoss(blank=0, reduction='mean'), ...)

Layout: `csrc/` kernels + C ABI + torch shim, `torch_asr/` the built native
modules (`torch_asr._ctc_lib`, named like the reference's `torch_asr._latgen_lib`),
`ctc/` the nn.Module / autograd surface, `decode.py` greedy decode + LER for `validate`,
`cabi.py` a ctypes view of the C ABI,
`synth.py` the synthetic workloads BASELINE.json names.

There is no CPU fallback: using the engine without its CUDA extension, or on a
non-CUDA tensor, raises.
"""
from .ctc import CTCLoss, ctc_loss, ctc_loss_parts  # noqa: F401

__all__ = ["CTCLoss", "ctc_loss", "ctc_loss_parts"]
# `pytorch_asr_b200.decode`: greedy decode + label error rate for `validate` (imported on demand)
