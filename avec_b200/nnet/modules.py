"""ConformerBlock sub-modules with the reference's constructor arguments, attribute names and state_dict keys
(reference nnet/modules.py:257-426, nnet/attentions.py:215-382, nnet/embeddings.py:101-158).  forward() dispatches to
the fused autograd Functions in avec_b200.functional (hand-written sm_100a kernels)."""
import torch
import torch.nn as nn

from .. import functional as AF
from .layers import Linear, Conv1d, Placeholder, Dropout, Swish

_pe_cache = {}
PE_MAX_LEN = 10000      # num_pos_embeddings / max_pos_encoding of every AVEC attention (networks.py:329)


def _drop_p(m, training):
    """effective dropout probability of a Dropout / Identity slot"""
    return float(m.p) if (training and isinstance(m, nn.Dropout)) else 0.0


def _pe_table(D, device, dtype, max_len=PE_MAX_LEN):
    """ONE sinusoid table per (D, device, dtype), as the reference keeps one buffer per module (embeddings.py:117-130):
    row r holds relative position max_len-1-r (even channels sin, odd cos), r = 0 .. 2*max_len-2.  Every length slices it, so
    device memory does not grow with the number of distinct sequence lengths seen in training.  Rows are pitched for TMA
    (ops.row_pitch).  The table is built on whichever stream asks first and then read from every stream (the audio and
    video encoders of the AV model run on two streams), hence the one-time synchronisation."""
    key = (D, str(device), dtype, max_len)
    t = _pe_cache.get(key)
    if t is None:
        pos = torch.arange(max_len - 1, -max_len, -1, dtype=torch.float).unsqueeze(1)
        angles = pos / 10000 ** (2 * torch.arange(0, D // 2, dtype=torch.float).unsqueeze(0) / D)
        pe = torch.zeros(2 * max_len - 1, D)
        pe[:, 0::2] = angles.sin()
        pe[:, 1::2] = angles.cos()
        from .. import ops
        t = ops.empty_rows(2 * max_len - 1, D, dtype, device)
        t.copy_(pe.to(device=device, dtype=dtype))
        if t.is_cuda and not torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream(t.device).synchronize()
        _pe_cache[key] = t
    return t


def rel_pos_table(T, D, device, dtype):
    """rows r = 0..2T-2 hold the sinusoid of relative position T-1-r; the slice pos_encoding[max_len-T : max_len-1+T] of
    RelativeSinusoidalPositionalEncoding (embeddings.py:117-152)."""
    assert T <= PE_MAX_LEN
    return _pe_table(D, device, dtype)[PE_MAX_LEN - T: PE_MAX_LEN - 1 + T]


def grouped_rel_pos_table(Tp, D, G, device, dtype):
    """GroupedRelativeSinusoidalPositionalEncoding slice for a (padded) length Tp = multiple of G, odd G
    (embeddings.py:160-216): 2*Tp - G rows, relative positions Tp-1-G//2 ... -(Tp-1-G//2)."""
    assert G % 2 == 1, "even group sizes use a different table layout in the reference and are not used by AVEC"
    half = Tp - 1 - G // 2
    assert half < PE_MAX_LEN
    return _pe_table(D, device, dtype)[PE_MAX_LEN - 1 - half: PE_MAX_LEN + half]


class FeedForwardModule(nn.Module):
    def __init__(self, dim_model, dim_ffn, drop_rate, act_fun="Swish", inner_dropout=True):
        super().__init__()
        assert act_fun == "Swish"
        self.layers = nn.Sequential(
            nn.LayerNorm(dim_model, eps=1e-6),
            Linear(dim_model, dim_ffn),
            Swish(),
            Dropout(p=drop_rate) if inner_dropout else Placeholder("Identity"),
            Linear(dim_ffn, dim_model),
            Dropout(p=drop_rate),
        )

    def forward_residual(self, x):
        """x + 1/2 * FFN(x): the half-step residual of ConformerBlock.forward (blocks.py:292,301) is fused in."""
        l = self.layers
        return AF.FFNFn.apply(x, l[0].weight, l[0].bias, l[1].weight, l[1].bias, l[4].weight, l[4].bias,
                              _drop_p(l[3], self.training), _drop_p(l[5], self.training))


class RelPos1dMultiHeadAttention(nn.Module):
    """Parameter holder of the relative-position attention (attentions.py:215-232); P = patch size (1 = regular)."""

    patch_size = 1

    def __init__(self, dim_model, num_heads, num_pos_embeddings=10000, attn_drop_rate=0.0, weight_init="default",
                 bias_init="default", output_proj=True, causal=False):
        super().__init__()
        assert not causal and output_proj and attn_drop_rate == 0.0
        self.num_heads = num_heads
        self.dim_model = dim_model
        self.dim_head = dim_model // num_heads
        self.max_len = num_pos_embeddings
        self.query_layer = Linear(dim_model, dim_model)
        self.key_layer = Linear(dim_model, dim_model)
        self.value_layer = Linear(dim_model, dim_model)
        self.output_layer = Linear(dim_model, dim_model)
        self.pos_layer = Linear(dim_model, dim_model)

    def forwardQKV(self, Q, K, V, mask=None):
        """the reference's plug-in entry point (attentions.py:280-323 / 348-382, called at modules.py:330): self-attention only
        (Q is K is V, as AttentionModule calls it); returns (output, None) - the attention map is never materialised."""
        return _forward_qkv(self, Q, K, V, mask)

    def forward(self, x, mask=None):
        return self.forwardQKV(x, x, x, mask)


class RelPosPatch1dMultiHeadAttention(RelPos1dMultiHeadAttention):
    def __init__(self, dim_model, num_heads, patch_size, num_pos_embeddings=10000, attn_drop_rate=0.0,
                 weight_init="default", bias_init="default", output_proj=True):
        super().__init__(dim_model, num_heads, num_pos_embeddings, attn_drop_rate, weight_init, bias_init, output_proj)
        self.patch_size = patch_size


class GroupedRelPosMultiHeadSelfAttention(nn.Module):
    """Parameter holder of the grouped / Transformer-XL relative attention (attentions.py:384-413, 556-577):
    u, v content / position biases registered before the projections, as in the reference's state_dict order."""

    def __init__(self, dim_model, num_heads, attn_drop_rate, max_pos_encoding, group_size, causal=False,
                 weight_init="scaled_uniform", bias_init="zeros", output_proj=True):
        super().__init__()
        assert not causal and output_proj and attn_drop_rate == 0.0 and (group_size * dim_model) % num_heads == 0
        self.num_heads, self.dim_model, self.group_size = num_heads, dim_model, group_size
        self.dim_head = (group_size * dim_model) // num_heads
        self.u = nn.Parameter(torch.zeros(dim_model))
        self.v = nn.Parameter(torch.zeros(dim_model))
        self.query_layer = Linear(dim_model, dim_model, bias_init=bias_init)
        self.key_layer = Linear(dim_model, dim_model, bias_init=bias_init)
        self.value_layer = Linear(dim_model, dim_model, bias_init=bias_init)
        self.output_layer = Linear(dim_model, dim_model, bias_init=bias_init)
        self.pos_layer = Linear(dim_model, dim_model)

    def forwardQKV(self, Q, K, V, mask=None):
        """attentions.py:579-650; self-attention only, returns (output, None)"""
        return _forward_qkv(self, Q, K, V, mask)

    def forward(self, x, mask=None):
        return self.forwardQKV(x, x, x, mask)


def _attention_args(a):
    return (a.query_layer.weight, a.query_layer.bias, a.key_layer.weight, a.key_layer.bias, a.value_layer.weight, a.value_layer.bias,
            a.output_layer.weight, a.output_layer.bias, a.pos_layer.weight, a.pos_layer.bias)


def _forward_qkv(a, Q, K, V, mask):
    """attention WITHOUT the module's LayerNorm / residual / dropout (the AttentionModule wraps those, modules.py:320-339):
    the fused Function is called with identity LayerNorm parameters switched off and the residual subtracted afterwards is
    avoided by the `plain` flag."""
    if not (Q is K and K is V):
        raise NotImplementedError("avec_b200: forwardQKV implements self-attention (Q is K is V), the only use in AVEC (modules.py:330)")
    from .blocks import mask_to_klen
    x = Q
    B, T, D = x.shape
    klen = mask_to_klen(mask).to(x.device) if mask is not None else None
    if isinstance(a, GroupedRelPosMultiHeadSelfAttention):
        G = a.group_size
        pe = grouped_rel_pos_table(-(-T // G) * G, D, G, x.device, x.dtype)
        y = AF.GroupedAttentionFn.apply(x, None, None, *_attention_args(a), a.u, a.v, pe, klen, a.num_heads, G, 0.0)
    else:
        P = a.patch_size
        pe = rel_pos_table(-(-T // P), D, x.device, x.dtype)
        y = AF.AttentionFn.apply(x, None, None, *_attention_args(a), pe, klen, a.num_heads, P, 0.0)
    return y, None


att_dict = {
    "RelPos1dMultiHeadAttention": RelPos1dMultiHeadAttention,
    "RelPosPatch1dMultiHeadAttention": RelPosPatch1dMultiHeadAttention,
    "GroupedRelPosMultiHeadSelfAttention": GroupedRelPosMultiHeadSelfAttention,
}


class AttentionModule(nn.Module):
    def __init__(self, dim_model, att_params, drop_rate, residual=False):
        super().__init__()
        if att_params["class"] not in att_dict:
            raise NotImplementedError(f"avec_b200: attention class {att_params['class']} is not implemented yet")
        self.norm = nn.LayerNorm(dim_model, eps=1e-6)
        self.attention = att_dict[att_params["class"]](dim_model=dim_model, **att_params["params"])
        self.dropout = Dropout(drop_rate)
        self.residual = residual

    def forward_residual(self, x, klen):
        """x + MHSA(LN(x)) with key-padding lengths klen (int32 [B] on device, or None)."""
        a = self.attention
        B, T, D = x.shape
        if isinstance(a, GroupedRelPosMultiHeadSelfAttention):
            G = a.group_size
            Tp = -(-T // G) * G
            pe = grouped_rel_pos_table(Tp, D, G, x.device, x.dtype)
            return AF.GroupedAttentionFn.apply(
                x, self.norm.weight, self.norm.bias, *_attention_args(a), a.u, a.v, pe, klen, a.num_heads, G,
                _drop_p(self.dropout, self.training))
        P = a.patch_size
        Tp = -(-T // P)
        pe = rel_pos_table(Tp, D, x.device, x.dtype)
        return AF.AttentionFn.apply(
            x, self.norm.weight, self.norm.bias, *_attention_args(a), pe, klen, a.num_heads, P, _drop_p(self.dropout, self.training))

    def forward(self, x, mask=None, hidden=None):
        """the reference's module call (modules.py:320-339): returns (x, attention map placeholder, hidden placeholder);
        with residual=False the caller adds the residual (blocks.py:295), so it is removed again here."""
        from .blocks import mask_to_klen
        y = self.forward_residual(x, mask_to_klen(mask).to(x.device) if mask is not None else None)
        return (y if self.residual else y - x), None, None


class ConvolutionModule(nn.Module):
    def __init__(self, dim_model, dim_expand, drop_rate, stride, act_fun="Swish", conv_params=None, channels_last=True,
                 batch_norm=True):
        super().__init__()
        assert act_fun == "Swish" and batch_norm and conv_params["class"] == "Conv1d"
        k = conv_params["params"]["kernel_size"]
        assert conv_params["params"].get("padding", "same") == "same" and k <= 15
        self.layers = nn.Sequential(
            nn.LayerNorm(dim_model, eps=1e-6),
            Conv1d(dim_model, 2 * dim_expand, kernel_size=1),
            Placeholder("GLU"),
            Conv1d(dim_expand, dim_expand, kernel_size=k, stride=stride, groups=dim_expand),
            nn.BatchNorm1d(dim_expand),
            Swish(),
            Conv1d(dim_expand, dim_expand, kernel_size=1),
            Dropout(p=drop_rate),
        )
        self.stride = stride

    def forward_residual(self, x, conv_res):
        """conv_res(x) + ConvModule(x) (blocks.py:298); conv_res is nn.Identity or a k=1 strided Conv1d."""
        l = self.layers
        bn = l[4]
        training = self.training
        if training and bn.track_running_stats:
            AF.count_batch(bn)
        has_res = not isinstance(conv_res, nn.Identity)
        return AF.ConvModuleFn.apply(
            x, l[0].weight, l[0].bias, l[1].weight, l[1].bias, l[3].weight, l[3].bias,
            bn.weight, bn.bias, bn.running_mean, bn.running_var, l[6].weight, l[6].bias,
            conv_res.weight if has_res else None, conv_res.bias if has_res else None,
            self.stride, training, bn.momentum, _drop_p(l[7], training))


class InterCTCResModule(nn.Module):
    def __init__(self, dim_model, vocab_size):
        super().__init__()
        self.proj_1 = Linear(dim_model, vocab_size)
        self.proj_2 = Linear(vocab_size, dim_model)

    def forward(self, x):
        return AF.InterCTCFn.apply(x, self.proj_1.weight, self.proj_1.bias, self.proj_2.weight, self.proj_2.bias)


class FusionModule(nn.Module):
    def __init__(self, a_dim_model=360, v_dim_model=360, f_dim_model=360, ff_ratio=4):
        super().__init__()
        self.layers = nn.Sequential(
            Linear(a_dim_model + v_dim_model, ff_ratio * f_dim_model),
            Swish(),
            Linear(ff_ratio * f_dim_model, f_dim_model),
        )

    def forward(self, audio, video):
        x = torch.cat([audio, video], dim=-1)
        l = self.layers
        return AF.MLPFn.apply(x, l[0].weight, l[0].bias, l[2].weight, l[2].bias)
