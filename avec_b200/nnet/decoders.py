"""Greedy CTC decoding with the reference's interface (nnet/decoders.py:75-120), on the device: argmax, length slicing,
repeat merging and blank removal are one kernel per batch (csrc/train.cu ctc_greedy_kernel); only the final token lists
cross to the host.  The frame-level argmax (`alignments`) is the "CTC alignment indices" output of BASELINE.json."""
import torch
import torch.nn as nn

from .. import ops


class CTCGreedySearchDecoder(nn.Module):
    def __init__(self, tokenizer_path=None, blank_token=0):
        super().__init__()
        self.tokenizer = None
        if tokenizer_path is not None:
            import sentencepiece as spm
            self.tokenizer = spm.SentencePieceProcessor(tokenizer_path)
        self.blank_token = blank_token

    def forward(self, outputs, from_logits=True):
        tokens = self.greedy_search(*outputs) if from_logits else outputs[0].tolist()
        return self.tokenizer.decode(tokens) if self.tokenizer is not None else tokens

    def greedy_search_device(self, logits, logits_len, want_align=False):
        """(tokens [B,T] int32 padded with -1, counts [B] int32[, alignments [B,T] int32]) - no host sync"""
        ln = logits_len.to(device=logits.device, dtype=torch.long) if logits_len is not None else None
        return ops.ctc_greedy_decode(logits.float().contiguous(), ln, self.blank_token, want_align=want_align)

    def greedy_search(self, logits, logits_len):
        tokens, ntok = self.greedy_search_device(logits, logits_len)
        tokens, ntok = tokens.cpu(), ntok.cpu()
        return [tokens[b, :int(ntok[b])].tolist() for b in range(tokens.shape[0])]
