"""CTC loss wrapper with the reference's call convention (nnet/losses.py:292-334): loss(targets, outputs) with
targets = (labels (B,L), label_lengths (B,)) and outputs = [logits (B,T,V), lengths (B,)]; mean over the batch of the
per-utterance negative log-likelihoods.  On CUDA tensors the loss and its gradient are one fused kernel (log-softmax +
alpha/beta recursions, csrc/ctc.cu) with device-side lengths, so a whole training step can be
captured in a CUDA graph."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CTCLoss(nn.Module):
    def __init__(self, blank=0, reduction="mean", zero_infinity=False, assert_shorter=True):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity, self.assert_shorter = blank, reduction, zero_infinity, assert_shorter

    def forward(self, targets, outputs):
        y, y_len = targets
        logits, logits_len = outputs
        if self.assert_shorter:   # host sync: keep off (assert_shorter=False) on the training hot path
            assert bool((y_len.cpu() <= logits_len.cpu()).all()), "ctc: label longer than logits"
        if logits.is_cuda:
            from .. import functional as AF
            loss = AF.CTCFn.apply(logits, y, logits_len, y_len, self.blank, self.zero_infinity)
            return loss.mean() if self.reduction == "mean" else loss.sum()
        logp = F.log_softmax(logits.float(), dim=-1).transpose(0, 1)
        # CPU length tensors are passed through untouched (no device sync: required under CUDA-graph capture)
        loss = F.ctc_loss(logp, y, logits_len.to(torch.long), y_len.to(torch.long), blank=self.blank, reduction="none",
                          zero_infinity=self.zero_infinity)
        return loss.mean() if self.reduction == "mean" else loss.sum()
