"""TEST INFRASTRUCTURE ONLY - CPU/GPU-agnostic fp32 restatement of the reference's hot-path algorithms in plain torch
functional ops, driven by a reference-format state_dict.  Pinned against fixtures produced by the unmodified reference
(tests/golden, oracle/make_golden.py) in tests/test_oracle.py; the `-m gpu` parity tests use it as the checker at
shapes the fixtures do not cover.  Never imported by the product package (avec_b200).

Every function cites the reference lines it follows (/root/reference/nnet/...).
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------- audio front-end
def mel_filterbank(n_freqs=257, f_min=0.0, f_max=8000.0, n_mels=80, sample_rate=16000):
    """torchaudio.functional.melscale_fbanks(257, 0, 8000, 80, 16000, norm=None, 'htk') (preprocessing.py:52)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(2595.0 * math.log10(1.0 + f_min / 700.0), 2595.0 * math.log10(1.0 + f_max / 700.0), n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    return torch.max(torch.zeros(1), torch.min(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


def logmel(wave):
    """AudioPreprocessing.forward (preprocessing.py:57-85): Spectrogram(512, 400, 160, hann periodic, center, reflect,
    power 2) -> MelScale(80) -> log(x + 1e-9).  (B, L) -> (B, 80, L // 160 + 1)."""
    B, Ln = wave.shape
    w = torch.zeros(512)
    w[56:456] = torch.hann_window(400, periodic=True)
    xp = F.pad(wave.unsqueeze(1), (256, 256), mode="reflect").squeeze(1)
    frames = xp.unfold(1, 512, 160)                       # (B, F, 512)
    spec = torch.fft.rfft(frames * w.to(wave), dim=-1)
    power = spec.real ** 2 + spec.imag ** 2               # (B, F, 257)
    mel = power @ mel_filterbank().to(wave)
    return torch.log(mel + 1e-9).transpose(1, 2)


# ---------------------------------------------------------------------------------------------------- ConformerBlock
def rel_pos_table(T, D):
    """RelativeSinusoidalPositionalEncoding slice [max_len-T : max_len-1+T] (embeddings.py:117-152)."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float).unsqueeze(1)
    angles = pos / 10000 ** (2 * torch.arange(0, D // 2, dtype=torch.float).unsqueeze(0) / D)
    pe = torch.zeros(2 * T - 1, D)
    pe[:, 0::2] = angles.sin()
    pe[:, 1::2] = angles.cos()
    return pe


def ffn(x, sd, p, m_in=None, m_out=None):
    """FeedForwardModule (modules.py:277-284).  m_in / m_out: the two nn.Dropout layers as explicit multipliers
    keep / (1 - p) (None = dropout off), so a test can feed the masks the kernels drew."""
    h = F.layer_norm(x, (x.shape[-1],), sd[p + "layers.0.weight"], sd[p + "layers.0.bias"], 1e-6)
    h = F.linear(h, sd[p + "layers.1.weight"], sd[p + "layers.1.bias"])
    h = h * torch.sigmoid(h)
    if m_in is not None:
        h = h * m_in
    h = F.linear(h, sd[p + "layers.4.weight"], sd[p + "layers.4.bias"])
    return h * m_out if m_out is not None else h


def relpos_attention(x, sd, p, klen, H, P):
    """AttentionModule (modules.py:320-339) + RelPos(Patch)1dMultiHeadAttention.forwardQKV (attentions.py:280-382).
    klen: (B,) valid key counts or None.  Index identity of rel_to_abs (attentions.py:258-276): out[i,j] = in[i, T-1+j-i]."""
    B, T, D = x.shape
    d = D // H
    h = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    a = p + "attention."
    Tq = T
    pad = (P - T % P) % P
    if pad:
        h = F.pad(h, (0, 0, 0, pad))
    Tp = (T + pad) // P
    if P > 1:
        h = h.view(B, Tp, P, D).mean(dim=2)                                     # AvgPool1d, padded zeros count
    q = F.linear(h, sd[a + "query_layer.weight"], sd[a + "query_layer.bias"]).view(B, Tp, H, d).transpose(1, 2)
    k = F.linear(h, sd[a + "key_layer.weight"], sd[a + "key_layer.bias"]).view(B, Tp, H, d).transpose(1, 2)
    v = F.linear(h, sd[a + "value_layer.weight"], sd[a + "value_layer.bias"]).view(B, Tp, H, d).transpose(1, 2)
    e = F.linear(rel_pos_table(Tp, D).to(x), sd[a + "pos_layer.weight"], sd[a + "pos_layer.bias"]).view(2 * Tp - 1, H, d).transpose(0, 1)
    s_k = q @ k.transpose(2, 3)
    s_e_rel = q @ e.transpose(1, 2).unsqueeze(0)                                 # (B,H,Tp,2Tp-1)
    idx = (Tp - 1) + torch.arange(Tp).unsqueeze(0) - torch.arange(Tp).unsqueeze(1)  # [i,j] -> T-1+j-i
    s_e = s_e_rel.gather(3, idx.to(x.device).expand(B, H, Tp, Tp))
    s = (s_k + s_e) / d ** 0.5
    # mask: key j valid iff every frame of its patch is < klen; query row valid iff its patch holds no padded frame
    keep = torch.ones(B, 1, Tp, Tp, device=x.device)
    if klen is not None:
        kl = torch.div(klen, P, rounding_mode="floor")
        keep = keep * (torch.arange(Tp, device=x.device)[None, None, None, :] < kl[:, None, None, None]).float()
    if P > 1:
        keep = keep * (torch.arange(Tp, device=x.device)[None, None, :, None] < (Tq // P)).float()
    if klen is not None or P > 1:
        s = s + (1.0 - keep) * -1e9
    w = s.softmax(dim=-1)
    o = (w @ v).transpose(1, 2).reshape(B, Tp, D)
    o = F.linear(o, sd[a + "output_layer.weight"], sd[a + "output_layer.bias"])
    if P > 1:
        o = o.repeat_interleave(P, dim=1)[:, :Tq]
    return o


def grouped_attention(x, sd, p, klen, H, G):
    """AttentionModule + GroupedRelPosMultiHeadSelfAttention.forwardQKV (attentions.py:579-650): projections at full
    length, zero-pad to a multiple of G AFTER the projections, G consecutive frames concatenated per token, Transformer-XL
    content / position biases u, v, grouped sinusoid table of 2*Tp-G rows (embeddings.py:160-216), mask[::G, ::G]."""
    B, T, D = x.shape
    h = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    a = p + "attention."
    q = F.linear(h, sd[a + "query_layer.weight"], sd[a + "query_layer.bias"])
    k = F.linear(h, sd[a + "key_layer.weight"], sd[a + "key_layer.bias"])
    v = F.linear(h, sd[a + "value_layer.weight"], sd[a + "value_layer.bias"])
    pad = (G - T % G) % G
    if pad:
        q, k, v = [F.pad(t, (0, 0, 0, pad)) for t in (q, k, v)]
    Tp = T + pad
    Tn, d = Tp // G, G * D // H
    qu, qv = q + sd[a + "u"], q + sd[a + "v"]
    half = Tp - 1 - G // 2
    pos = torch.arange(half, -half - 1, -1, dtype=torch.float).unsqueeze(1)
    ang = pos / 10000 ** (2 * torch.arange(0, D // 2, dtype=torch.float).unsqueeze(0) / D)
    pe = torch.zeros(2 * Tp - G, D)
    pe[:, 0::2], pe[:, 1::2] = ang.sin(), ang.cos()
    e = F.linear(pe.to(x), sd[a + "pos_layer.weight"], sd[a + "pos_layer.bias"])
    qu, qv, k, v = [t.reshape(B, Tn, H, d).transpose(1, 2) for t in (qu, qv, k, v)]
    e = e.reshape(2 * Tn - 1, H, d).transpose(0, 1)
    idx = (Tn - 1) + torch.arange(Tn).unsqueeze(0) - torch.arange(Tn).unsqueeze(1)
    s = (qu @ k.transpose(2, 3) + (qv @ e.transpose(1, 2).unsqueeze(0)).gather(3, idx.to(x.device).expand(B, H, Tn, Tn))) / d ** 0.5
    if klen is not None:
        kl = torch.div(klen + G - 1, G, rounding_mode="floor")
        keep = (torch.arange(Tn, device=x.device)[None, None, None, :] < kl[:, None, None, None]).float()
        s = s + (1.0 - keep) * -1e9
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, Tp, D)[:, :T]
    return F.linear(o, sd[a + "output_layer.weight"], sd[a + "output_layer.bias"])


def conv_module(x, sd, p, stride, training, bn_momentum=0.1):
    """ConvolutionModule (modules.py:372-381): LN -> PW conv -> GLU -> depthwise k15 (same pad 7/7) -> BN1d -> Swish -> PW."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[p + "layers.0.weight"], sd[p + "layers.0.bias"], 1e-6).transpose(1, 2)
    h = F.conv1d(h, sd[p + "layers.1.weight"], sd[p + "layers.1.bias"])
    h = F.glu(h, dim=1)
    wd = sd[p + "layers.3.weight"]
    k = wd.shape[-1]
    h = F.conv1d(F.pad(h, ((k - 1) // 2, k // 2)), wd, sd[p + "layers.3.bias"], stride=stride, groups=wd.shape[0])
    rm, rv = sd[p + "layers.4.running_mean"].clone(), sd[p + "layers.4.running_var"].clone()
    h = F.batch_norm(h, rm, rv, sd[p + "layers.4.weight"], sd[p + "layers.4.bias"], training, bn_momentum, 1e-5)
    h = h * torch.sigmoid(h)
    h = F.conv1d(h, sd[p + "layers.6.weight"], sd[p + "layers.6.bias"])
    return h.transpose(1, 2), rm, rv


def conformer_block(x, sd, klen, H, P, stride, training=True, prefix="", G=None, drop=None):
    """ConformerBlock.forward (blocks.py:289-306).  G: group size when the block uses grouped attention.
    drop: None (all dropouts off) or the six dropout multipliers of the block in call order
    (ff1 inner, ff1 out, attention out, conv out, ff2 inner, ff2 out; modules.py:281,283,333,380)."""
    p = prefix
    d = drop if drop is not None else [None] * 6
    x = x + 0.5 * ffn(x, sd, p + "ff_module1.", d[0], d[1])
    if G is not None:
        a = grouped_attention(x, sd, p + "self_att_module.", klen, H, G)
    else:
        a = relpos_attention(x, sd, p + "self_att_module.", klen, H, P)
    x = x + (a * d[2] if d[2] is not None else a)
    c, rm, rv = conv_module(x, sd, p + "conv_module.", stride, training)
    if d[3] is not None:
        c = c * d[3]
    if p + "conv_res.weight" in sd:
        r = F.conv1d(x.transpose(1, 2), sd[p + "conv_res.weight"], sd[p + "conv_res.bias"], stride=stride).transpose(1, 2)
    else:
        r = x
    x = r + c
    x = x + 0.5 * ffn(x, sd, p + "ff_module2.", d[4], d[5])
    De = x.shape[-1]
    return F.layer_norm(x, (De,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6), rm, rv


def interctc(x, sd, p):
    """InterCTCResModule (modules.py:394-400)."""
    logits = F.linear(x, sd[p + "proj_1.weight"], sd[p + "proj_1.bias"])
    return x + F.linear(logits.softmax(dim=-1), sd[p + "proj_2.weight"], sd[p + "proj_2.bias"]), logits


# ------------------------------------------------------------------------------------------------------------ ResNet
def _bn2d(h, sd, p, training):
    return F.batch_norm(h, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd[p + "weight"], sd[p + "bias"],
                        training, 0.1, 1e-5)


def resnet_block(x, sd, stride, training=True, prefix=""):
    """ResNetBlock.forward (blocks.py:64-91) on NCHW input; 'same' pre-padding (layers.py:250-258) then valid conv."""
    p = prefix
    h = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[p + "layers.0.weight"], None, stride=stride)
    h = F.relu(_bn2d(h, sd, p + "layers.1.", training))
    h = F.conv2d(F.pad(h, (1, 1, 1, 1)), sd[p + "layers.3.weight"], None)
    h = _bn2d(h, sd, p + "layers.4.", training)
    if p + "residual.0.weight" in sd:
        r = _bn2d(F.conv2d(x, sd[p + "residual.0.weight"], None, stride=stride), sd, p + "residual.1.", training)
    else:
        r = x
    return F.relu(h + r)


def video_stem(video, sd, training=True, prefix=""):
    """Conv3d(1->64,(5,7,7),s(1,2,2),same)+BN3d+ReLU -> MaxPool3d((1,3,3),s(1,2,2), zero same pad) -> VideoToImages
    (networks.py:459-471, layers.py:839-915, transforms.py:68-72).  video (B,1,T,H,W) -> (B*T, 64, H/4, W/4)."""
    p = prefix
    h = F.conv3d(F.pad(video, (3, 3, 3, 3, 2, 2)), sd[p + "0.layers.0.0.weight"], sd[p + "0.layers.0.0.bias"], stride=(1, 2, 2))
    h = F.batch_norm(h, sd[p + "0.layers.0.1.running_mean"].clone(), sd[p + "0.layers.0.1.running_var"].clone(),
                     sd[p + "0.layers.0.1.weight"], sd[p + "0.layers.0.1.bias"], training, 0.1, 1e-5)
    h = F.relu(h)
    h = F.max_pool3d(F.pad(h, (1, 1, 1, 1, 0, 0)), (1, 3, 3), (1, 2, 2))
    return h.transpose(1, 2).flatten(0, 1)


def audio_stem(mel, sd, training=True, prefix=""):
    """unsqueeze -> Conv2d(1->180,k3,s2,same)+BN2d+Swish -> reshape (B, C*F', T') -> transpose -> Linear(7200->180)
    (networks.py:359-377, 420-432).  mel (B,80,T) -> (B,T',180)."""
    p = prefix
    h = F.conv2d(F.pad(mel.unsqueeze(1), (1, 1, 1, 1)), sd[p + "subsampling_module.layers.0.0.weight"],
                 sd[p + "subsampling_module.layers.0.0.bias"], stride=2)
    h = F.batch_norm(h, sd[p + "subsampling_module.layers.0.1.running_mean"].clone(),
                     sd[p + "subsampling_module.layers.0.1.running_var"].clone(), sd[p + "subsampling_module.layers.0.1.weight"],
                     sd[p + "subsampling_module.layers.0.1.bias"], training, 0.1, 1e-5)
    h = h * torch.sigmoid(h)
    B, C, Fq, T = h.shape
    h = h.reshape(B, C * Fq, T).transpose(1, 2)
    return F.linear(h, sd[p + "linear.weight"], sd[p + "linear.bias"])


# ------------------------------------------------------------------------------------------------------ full encoders
def conformer_stack(x, lengths, sd, p, dims, num_blocks, interctc_blocks, patch, loss_prefix, training=True, H=4):
    """ConformerInterCTC.forward (networks.py:262-307).  dims/num_blocks per stage, patch = patch size per stage."""
    outs = {}
    klen = lengths
    i, j = 0, 0
    for stage, nb in enumerate(num_blocks):
        for b in range(nb):
            down = (b == nb - 1) and (stage < len(num_blocks) - 1)
            stride = 2 if down else 1
            x, _, _ = conformer_block(x, sd, klen, H, patch[stage], stride, training, prefix=f"{p}conformer_blocks.{i}.")
            logits = None
            if i + 1 in interctc_blocks:
                x, logits = interctc(x, sd, f"{p}interctc_modules.{j}.")
                j += 1
            if stride > 1:
                lengths = torch.div(lengths - 1, stride, rounding_mode="floor") + 1
                klen = lengths
            if logits is not None:
                outs[f"{loss_prefix}_{i}"] = [logits, lengths]
            i += 1
    return x, lengths, outs


def visual_encoder(video, vlen, sd, p, num_blocks, interctc_blocks, loss_prefix, training=True, head=True):
    """VisualEfficientConformerEncoder.forward (networks.py:497-512); video (B,T,H,W,1)."""
    B, T = video.shape[0], video.shape[1]
    h = video_stem(video.permute(0, 4, 1, 2, 3), sd, training, p + "front_end.")
    strides = [1, 1, 2, 1, 2, 1, 2, 1]
    for k in range(8):
        h = resnet_block(h, sd, strides[k], training, f"{p}front_end.3.blocks.{k}.")
    h = F.linear(h.mean((2, 3)), sd[p + "front_end.3.head.1.weight"], sd[p + "front_end.3.head.1.bias"]).view(B, T, -1)
    x, lengths, outs = conformer_stack(h, vlen, sd, p + "back_end.", [256, 360], num_blocks, interctc_blocks, [1, 1], loss_prefix, training)
    if head:
        x = F.linear(x, sd[p + "head.weight"], sd[p + "head.bias"])
    return x, lengths, outs


def audio_encoder(audio, alen, sd, p, num_blocks, interctc_blocks, loss_prefix, training=True, head=True, att_patch=3):
    """AudioEfficientConformerEncoder.forward (networks.py:411-440), SpecAugment bypassed."""
    mel = logmel(audio)
    lengths = torch.div(alen, 160, rounding_mode="floor") + 1
    x = audio_stem(mel, sd, training, p)
    lengths = torch.div(lengths - 1, 2, rounding_mode="floor") + 1
    x, lengths, outs = conformer_stack(x, lengths, sd, p + "back_end.", [180, 256, 360], num_blocks, interctc_blocks,
                                       [att_patch, 1, 1], loss_prefix, training)
    if head:
        x = F.linear(x, sd[p + "head.weight"], sd[p + "head.bias"])
    return x, lengths, outs


def av_model(sd, video, vlen, audio, alen, training=True):
    """AudioVisualEfficientConformerInterCTC.forward (models_zoo.py:156-161, networks.py:559-579), default InterCTC blocks."""
    p = "encoder."
    v, vl, vo = visual_encoder(video, vlen, sd, p + "video_encoder.", [6, 1], [3, 6], "v_ctc", training, head=False)
    a, al, ao = audio_encoder(audio, alen, sd, p + "audio_encoder.", [5, 6, 1], [8, 11], "a_ctc", training, head=False)
    x = torch.cat([a, v], dim=-1)
    x = F.linear(x, sd[p + "fusion_module.layers.0.weight"], sd[p + "fusion_module.layers.0.bias"])
    x = x * torch.sigmoid(x)
    x = F.linear(x, sd[p + "fusion_module.layers.2.weight"], sd[p + "fusion_module.layers.2.bias"])
    x, lengths, fo = conformer_stack(x, al, sd, p + "audio_visual_encoder.", [360], [5], [2], [1], "f_ctc", training)
    x = F.linear(x, sd[p + "head.weight"], sd[p + "head.bias"])
    outs = {"outputs": [x, lengths]}
    outs.update(fo)
    outs.update(vo)
    outs.update(ao)
    return outs


def ao_model(sd, audio, alen, training=True, interctc_blocks=(3, 6, 10, 13)):
    x, lengths, outs = audio_encoder(audio, alen, sd, "encoder.", [5, 6, 5], list(interctc_blocks), "ctc", training)
    o = {"outputs": [x, lengths]}
    o.update(outs)
    return o


def vo_model(sd, video, vlen, training=True, interctc_blocks=(3, 6, 9)):
    x, lengths, outs = visual_encoder(video, vlen, sd, "encoder.", [6, 6], list(interctc_blocks), "ctc", training)
    o = {"outputs": [x, lengths]}
    o.update(outs)
    return o
