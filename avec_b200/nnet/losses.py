"""CTC loss wrapper with the reference's call convention (nnet/losses.py:292-334): loss(targets, outputs) with
targets = (labels (B,L), label_lengths (B,)) and outputs = [logits (B,T,V), lengths (B,)]; mean over the batch of the
per-utterance negative log-likelihoods.  On CUDA tensors the loss and its gradient are one fused kernel (log-softmax +
alpha/beta recursions, csrc/ctc.cu) with device-side lengths, so a whole training step can be
captured in a CUDA graph."""
import torch
import torch.nn as nn


class CTCLoss(nn.Module):
    def __init__(self, blank=0, reduction="mean", zero_infinity=False, assert_shorter=True):
        super().__init__()
        # the reference's "mean" is the mean over the batch of the per-utterance losses (losses.py:331-333); "sum" their sum
        assert reduction in ("mean", "sum"), f"CTCLoss reduction {reduction!r}: the reference implements 'mean' and 'sum'"
        self.blank, self.reduction, self.zero_infinity, self.assert_shorter = blank, reduction, zero_infinity, assert_shorter

    def forward(self, targets, outputs):
        y, y_len = targets
        logits, logits_len = outputs
        if not logits.is_cuda:
            raise RuntimeError("avec_b200.nnet.CTCLoss needs CUDA logits: the hot path has no CPU fallback")
        if self.assert_shorter:   # host sync: keep off (assert_shorter=False) on the training hot path
            assert bool((y_len.cpu() <= logits_len.cpu()).all()), "ctc: label longer than logits"
        from .. import functional as AF
        loss = AF.CTCFn.apply(logits, y, logits_len, y_len, self.blank, self.zero_infinity)
        return loss.mean() if self.reduction == "mean" else loss.sum()
