"""AudioPreprocessing with the reference's interface (nnet/preprocessing.py:24-85): (B, L) waveform ->
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
