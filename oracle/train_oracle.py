"""TEST INFRASTRUCTURE ONLY - numpy / torch-CPU restatement of the training-step pieces around the encoder hot path
(SURVEY section 8(f) rows 2-4).  Never imported by the product package (avec_b200).

* philox4x32_10 / dropout_keep: the counter-based generator of csrc/train.cu (Philox4x32-10, Salmon et al. SC'11; the
  published known-answer vectors of the Random123 distribution are checked in tests/test_train_oracle.py), so the GPU
  dropout masks and SpecAugment intervals are compared BIT-EXACTLY, and block parity with dropout on is tested by feeding
  these masks to oracle/restate.py.
* spec_augment: nnet/preprocessing.py:118-127 on top of torchaudio.functional.mask_along_axis (torchaudio 2.11:
  value = rand * mask_param; min_value = rand * (size - value); zeros in [int(min_value), int(min_value) + int(value))).
* greedy_decode: nnet/decoders.py:97-120.
* adam_step: torch.optim.Adam (weight_decay = L2) as subclassed in nnet/optimizers.py:61-93, NoamDecayScheduler
  (nnet/schedulers.py:120-137), torch.nn.utils.clip_grad_norm_ (nnet/model.py:378-380), EMA (nnet/model.py:401-404).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10: counters are uint32 arrays (broadcastable), key two python ints.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def dropout_keep(seed, step, site, rows, C, p):
    """bool [rows, C]: element (r, c) kept iff its 16 bits >= round(p * 65536) (csrc/train.cu dropout_kernel)."""
    groups = (C + 7) // 8
    r = np.arange(rows, dtype=np.uint64)[:, None]
    g = np.arange(groups, dtype=np.uint64)[None, :]
    w = philox4x32_10(r, g, np.uint64(site), np.uint64(step & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.empty((rows, groups, 8), dtype=np.uint32)
    for j in range(8):
        word = w[j >> 1]
        bits[:, :, j] = (word >> np.uint32(16)) if (j & 1) else (word & np.uint32(0xFFFF))
    thresh = int(np.float32(p) * np.float32(65536.0) + np.float32(0.5))
    return (bits.reshape(rows, groups * 8)[:, :C] >= thresh)


def dropout_scale_mask(seed, step, site, rows, C, p):
    """float32 multiplier keep / (1 - p), as nn.Dropout scales the survivors"""
    return dropout_keep(seed, step, site, rows, C, p).astype(np.float32) * (np.float32(1.0) / (np.float32(1.0) - np.float32(p)))


def _u01(w):
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def _interval(seed, step, site, b, k, param, size):
    w = philox4x32_10(np.uint64(b), np.uint64(k), np.uint64(site), np.uint64(step & 0xFFFFFFFF), seed & 0xFFFFFFFF,
                      (seed >> 32) & 0xFFFFFFFF)
    value = _u01(w[0]) * np.float32(param)
    minv = _u01(w[1]) * (np.float32(size) - value)
    lo = int(minv)
    return lo, lo + int(value)


def spec_augment_intervals(seed, step, site, lengths, F, M, mF=2, Fmax=27, mT=5, pS=0.05):
    """[B][mF+mT] (lo, hi) zeroed intervals: the first mF over mel bins (shared by the batch), the rest over frames."""
    out = []
    for b, ln in enumerate(lengths):
        ln = min(int(ln), F)
        iv = []
        for k in range(mF):
            param = min(Fmax, M)
            iv.append(_interval(seed, step, site, 0xFFFFFFFF, k, param, M) if param >= 1 else (0, 0))
        Tb = min(int(np.float32(pS) * np.float32(ln)), ln)
        for k in range(mT):
            iv.append(_interval(seed, step, site, b, k, Tb, ln) if Tb >= 1 else (0, 0))
        out.append(iv)
    return out


def spec_augment(mel, lengths, seed, step, site, mF=2, Fmax=27, mT=5, pS=0.05):
    """mel [B, F, M] (frame-major) numpy float32 -> masked copy (SpecAugment.forward, training branch)."""
    B, F, M = mel.shape
    out = mel.copy()
    ivs = spec_augment_intervals(seed, step, site, lengths, F, M, mF, Fmax, mT, pS)
    for b in range(B):
        for k, (lo, hi) in enumerate(ivs[b]):
            if k < mF:
                out[b, :, lo:hi] = 0.0
            else:
                ln = min(int(lengths[b]), F)
                out[b, lo:min(hi, ln), :] = 0.0
    return out


def video_draws(seed, step, site, b, length, Hi, Wi, Ho, Wo, flip_p=0.5, mask_T=10, fps=25.0, num_mask_second=1.0):
    """(crop row, crop column, flip, [(lo, hi) per time mask]) of sample b: csrc/train.cu video_draw / mask_interval"""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    w = philox4x32_10(np.uint64(b), np.uint64(0), np.uint64(site), np.uint64(step & 0xFFFFFFFF), k0, k1)
    oy = min(int(_u01(w[0]) * np.float32(Hi - Ho + 1)), Hi - Ho)
    ox = min(int(_u01(w[1]) * np.float32(Wi - Wo + 1)), Wi - Wo)
    flip = bool(_u01(w[2]) < np.float32(flip_p))
    nmask = int(np.float32(length) / np.float32(fps) * np.float32(num_mask_second))
    return oy, ox, flip, [_interval(seed, step, site, b, 1 + m, mask_T, length) for m in range(nmask)]


def video_augment(video, lengths, seed, step, site, crop=(88, 88), flip_p=0.5, mask_T=10, fps=25.0, num_mask_second=1.0):
    """video [B,T,Hi,Wi] float32 numpy -> [B,T,Ho,Wo]: torchvision RandomCrop -> RandomHorizontalFlip -> nnet.TimeMaskSecond
    (transforms.py:108-126, mean_frame=True) per sample on its first lengths[b] frames, zeros beyond (collate padding)."""
    B, T, Hi, Wi = video.shape
    Ho, Wo = crop
    out = np.zeros((B, T, Ho, Wo), dtype=np.float32)
    for b in range(B):
        ln = min(int(lengths[b]), T)
        oy, ox, flip, masks = video_draws(seed, step, site, b, ln, Hi, Wi, Ho, Wo, flip_p, mask_T, fps, num_mask_second)
        clip = video[b, :ln, oy:oy + Ho, ox:ox + Wo].copy()
        if flip:
            clip = clip[:, :, ::-1].copy()
        for lo, hi in masks:
            clip[lo:min(hi, ln)] = np.float32(clip.astype(np.float64).mean())
        out[b, :ln] = clip
    return out


def greedy_decode(logits, lengths, blank=0):
    """CTCGreedySearchDecoder.greedy_search: list of token lists; logits [B,T,V] numpy, lengths [B]."""
    preds = logits.argmax(axis=-1)
    out = []
    for b in range(logits.shape[0]):
        seq, prev = [], None
        for t in range(int(lengths[b])):
            tok = int(preds[b, t])
            if tok != prev and tok != blank:
                seq.append(tok)
            prev = tok
        out.append(seq)
    return out


def noam_lr(step, warmup_steps=10000, dim_decay=360, val_factor=2):
    return val_factor * dim_decay ** -0.5 * min(step * warmup_steps ** -1.5, step ** -0.5)


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.98, eps=1e-9, wd=1e-6, max_norm=None, ema=None, tau=0.0):
    """one torch.optim.Adam step (float64 arithmetic on numpy arrays, returns new p, m, v, ema)"""
    p, g, m, v = [np.asarray(t, dtype=np.float64) for t in (p, g, m, v)]
    if max_norm is not None:
        g = g * min(1.0, max_norm / (np.sqrt((g ** 2).sum()) + 1e-6))
    g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - lr / (1 - b1 ** step) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** step) + eps)
    if ema is not None:
        ema = tau * np.asarray(ema, dtype=np.float64) + (1 - tau) * p
    return p, m, v, ema
