"""Deterministic parameters / inputs shared by oracle/make_golden.py (reference side, CPU) and the GPU parity tests.
Values depend only on (parameter name, shape, seed), never on module construction order, so both sides agree."""
import hashlib
import math

import torch


def _gen(name, seed):
    h = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:8], "little") % (2 ** 63 - 1)
    return torch.Generator().manual_seed(h)


def seeded_state_dict(model, seed=0):
    """returns {key: tensor} for every state_dict entry of `model` (buffers of the audio transforms are left alone)."""
    out = {}
    for k, v in model.state_dict().items():
        g = _gen(k, seed)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("Spectrogram.window") or k.endswith("MelScale.fb"):
            out[k] = v.clone()
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif v.dim() == 1 and k.endswith("weight"):      # LayerNorm / BatchNorm gains
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:                                 # biases, u/v
            out[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(fan_in)
    return out


def randn(name, shape, seed=0, scale=1.0):
    return scale * torch.randn(shape, generator=_gen("input:" + name, seed))


def subsample(t, max_n=4096):
    """fixed strided subsample of a tensor + its sum and L2 norm: compact fingerprint of a large gradient."""
    f = t.detach().float().flatten()
    stride = max(1, -(-f.numel() // max_n))
    return {"stride": stride, "vals": f[::stride].clone(), "sum": float(f.double().sum()), "norm": float(f.double().norm())}
