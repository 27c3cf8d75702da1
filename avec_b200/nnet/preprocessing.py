"""AudioPreprocessing / SpecAugment with the reference's interface (nnet/preprocessing.py:24-85): (B, L) waveform ->
(B, 80, L // 160 + 1) log-mel, lengths // 160 + 1, computed by the fused sm_100a STFT->mel->log kernel."""
import torch
import torch.nn as nn

from .. import ops
from .layers import mel_filterbank


class AudioPreprocessing(nn.Module):
    def __init__(self, sample_rate=16000, n_fft=512, win_length_ms=25, hop_length_ms=10, n_mels=80, normalize=False, mean=0, std=1):
        super().__init__()
        assert (sample_rate, n_fft, win_length_ms, hop_length_ms, n_mels) == (16000, 512, 25, 10, 80) and not normalize
        self.hop_length = 160
        self.Spectrogram = nn.Module()
        self.Spectrogram.register_buffer("window", torch.hann_window(400))
        self.MelScale = nn.Module()
        self.MelScale.register_buffer("fb", mel_filterbank())

    def forward(self, x, lengths=None):
        out = ops.stft_mel_log(x.float().contiguous(), self.MelScale.fb, layout=1).to(x.dtype)
        if lengths is not None:
            lengths = torch.div(lengths, self.hop_length, rounding_mode="floor") + 1
            return out, lengths
        return out


class SpecAugment(nn.Module):
    """SpecAugment(mF, F, mT, pS) of the reference (nnet/preprocessing.py:87-129), on the device: mF frequency masks shared by
    the batch, mT time masks per utterance of width < int(pS * length), drawn by the counter-based generator of csrc/train.cu
    (no per-sample Python loop, no host sync; capturable in a CUDA graph).  Inside the audio encoder the kernel runs in the
    fused stem between the log-mel kernel and the stem convolution; called directly it takes the reference's (B, n_mels, T)."""

    def __init__(self, mF=2, F=27, mT=5, pS=0.05):
        super().__init__()
        self.mF, self.F, self.mT, self.pS = mF, F, mT, pS
        self.enabled = True          # nnet.zero_dropout() clears it (deterministic parity configuration)

    def params(self):
        return (self.mF, self.F, self.mT, self.pS)

    def forward(self, samples, lengths):
        if not (self.training and self.enabled):
            return samples
        mel = samples.float().transpose(1, 2).contiguous()                     # [B, T, n_mels] frame-major
        ln = lengths.to(device=mel.device, dtype=torch.long) if lengths is not None else None
        ops.spec_augment_(mel, ln, ops.RNG.next_site(), *self.params())
        return mel.transpose(1, 2).to(samples.dtype)

    def extra_repr(self):
        return f"mF={self.mF}, F={self.F}, mT={self.mT}, pS={self.pS}"
