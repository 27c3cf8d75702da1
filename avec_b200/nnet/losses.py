"""CTC loss wrapper with the reference's call convention (nnet/losses.py:292-334): loss(targets, outputs) with
targets = (labels (B,L), label_lengths (B,)) and outputs = [logits (B,T,V), lengths (B,)]; mean over the batch of the
per-utterance negative log-likelihoods.  SURVEY section 8(f) ranks a fused CTC kernel as the *next* row; until then the
loss itself (not on the encoder hot path) is torch's log_softmax + ctc_loss."""
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
        if self.assert_shorter:
            assert bool((y_len <= logits_len.to(y_len.device)).all()), "ctc: label longer than logits"
        logp = F.log_softmax(logits.float(), dim=-1).transpose(0, 1)
        # CPU length tensors are passed through untouched (no device sync: required under CUDA-graph capture)
        loss = F.ctc_loss(logp, y, logits_len.to(torch.long), y_len.to(torch.long), blank=self.blank, reduction="none",
                          zero_infinity=self.zero_infinity)
        return loss.mean() if self.reduction == "mean" else loss.sum()
