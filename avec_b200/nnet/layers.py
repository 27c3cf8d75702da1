"""Parameter-holding primitives with the reference's names, shapes and initialisation (nnet/layers.py:29-503,
nnet/normalizations.py:27-170, nnet/activations.py:39-45 of the reference).  Their arithmetic lives in the fused
Functions of avec_b200.functional; these classes only own the parameters so that state_dict() matches SURVEY A.2."""
import math

import torch
import torch.nn as nn
import torch.nn.init as init


def _he_normal_(t):
    return init.kaiming_normal_(t)


class Linear(nn.Linear):
    def __init__(self, in_features, out_features, bias=True, weight_init="default", bias_init="default"):
        super().__init__(in_features, out_features, bias=bias)
        if weight_init == "he_normal":
            _he_normal_(self.weight)
        if bias_init == "zeros" and self.bias is not None:
            init.zeros_(self.bias)


class Conv1d(nn.Conv1d):
    """k-tap 1-d convolution parameters, weight (Cout, Cin/groups, k) (checkpoint layout)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, groups=1, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, groups=groups, bias=bias, padding=0)


class Conv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=True, weight_init="default"):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, bias=bias, padding=0)
        if weight_init == "he_normal":
            _he_normal_(self.weight)


class Conv3d(nn.Conv3d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, bias=bias, padding=0)


class Swish(nn.Module):
    """x * sigmoid(x) - evaluated inside the fused kernels (GEMM epilogue / BN-apply)."""

    def forward(self, x):  # pragma: no cover - never called on the hot path
        raise RuntimeError("avec_b200: Swish is fused into the producing kernel; call the parent module instead")


class Placeholder(nn.Module):
    """Keeps nn.Sequential indices aligned with the reference (GLU, ReLU, Identity, Dropout slots)."""

    def __init__(self, what=""):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class Dropout(nn.Dropout):
    """Holder of the reference's p.  The owning module applies it inside its fused Function (counter-based Philox masks
    regenerated in the backward, csrc/train.cu avec_dropout); calling this layer directly dispatches to the same kernel."""

    def forward(self, x):
        if not self.training or self.p == 0:
            return x
        from .. import functional as AF
        return AF.DropoutFn.apply(x, float(self.p))


def mel_filterbank(n_freqs=257, f_min=0.0, f_max=8000.0, n_mels=80, sample_rate=16000):
    """HTK mel filterbank, norm=None: the matrix torchaudio.functional.melscale_fbanks builds for
    MelScale(80, 16000, 0, 8000, n_stft=257) (reference nnet/preprocessing.py:52)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))
