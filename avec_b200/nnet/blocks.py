"""ConformerBlock and ResNetBlock (reference nnet/blocks.py:29-91, 208-306) on the fused sm_100a kernels."""
import torch
import torch.nn as nn

from .. import functional as AF
from .layers import Conv1d, Conv2d, Placeholder
from .modules import FeedForwardModule, AttentionModule, ConvolutionModule


def mask_to_klen(mask):
    """reference-style float mask (B,1,T,T) (1 = keep; key-padding prefix of ones, attentions.py:682-692) -> int32 [B]."""
    if mask is None:
        return None
    return mask[:, 0, 0, :].sum(dim=-1).to(torch.int32)


class ConformerBlock(nn.Module):
    def __init__(self, dim_model, dim_expand, ff_ratio, att_params, drop_rate, conv_stride, conv_params, inner_dropout=True,
                 act_fun="Swish", batch_norm=True, block_norm=True):
        super().__init__()
        assert block_norm
        self.ff_module1 = FeedForwardModule(dim_model, dim_model * ff_ratio, drop_rate, act_fun, inner_dropout)
        self.self_att_module = AttentionModule(dim_model, att_params, drop_rate, residual=False)
        self.conv_module = ConvolutionModule(dim_model, dim_expand, drop_rate, conv_stride, act_fun, conv_params,
                                             channels_last=True, batch_norm=batch_norm)
        self.ff_module2 = FeedForwardModule(dim_expand, dim_expand * ff_ratio, drop_rate, act_fun, inner_dropout)
        self.norm = nn.LayerNorm(dim_expand, eps=1e-6)
        if dim_model != dim_expand:
            self.conv_res = Conv1d(dim_model, dim_expand, kernel_size=1, stride=conv_stride)
        elif conv_stride > 1:
            # the reference's layer_dict has no "MaxPool1d" entry, so this branch raises KeyError there
            raise KeyError("MaxPool1d")
        else:
            self.conv_res = nn.Identity()
        self.stride = conv_stride

    def forward(self, x, mask=None, klen=None):
        """x (B,T,D) in the compute dtype -> (B,T',De).  `mask` may be the reference's float mask; the fused attention
        kernel only needs the per-item number of valid keys (`klen`)."""
        if klen is None and mask is not None:
            klen = mask_to_klen(mask)
        x = self.ff_module1.forward_residual(x)
        x = self.self_att_module.forward_residual(x, klen)
        x = self.conv_module.forward_residual(x, self.conv_res)
        x = self.ff_module2.forward_residual(x)
        return AF.LayerNormFn.apply(x, self.norm.weight, self.norm.bias)


class ResNetBlock(nn.Module):
    """BasicBlock on channels-last images [N,H,W,C]."""

    def __init__(self, in_features, out_features, kernel_size=(3, 3), stride=(1, 1), act_fun="ReLU", joined_post_act=True):
        super().__init__()
        assert tuple(kernel_size) == (3, 3) and act_fun == "ReLU" and joined_post_act
        s = stride[0] if isinstance(stride, (tuple, list)) else stride
        self.layers = nn.Sequential(
            Conv2d(in_features, out_features, (3, 3), stride=(s, s), bias=False, weight_init="he_normal"),
            nn.BatchNorm2d(out_features),
            Placeholder("ReLU"),
            Conv2d(out_features, out_features, (3, 3), bias=False, weight_init="he_normal"),
            nn.BatchNorm2d(out_features),
            Placeholder("Identity"),
        )
        self.joined_post_act = Placeholder("ReLU")
        if s > 1 or in_features != out_features:
            self.residual = nn.Sequential(
                Conv2d(in_features, out_features, 1, stride=(s, s), bias=False, weight_init="he_normal"),
                nn.BatchNorm2d(out_features),
            )
        else:
            self.residual = nn.Identity()
        self.stride = s

    def forward(self, x):
        l = self.layers
        bn1, bn2 = l[1], l[4]
        has_res = not isinstance(self.residual, nn.Identity)
        training = self.training
        if training:
            AF.count_batch(bn1)
            AF.count_batch(bn2)
            if has_res:
                AF.count_batch(self.residual[1])
        if has_res:
            cr, br = self.residual[0], self.residual[1]
            res = (cr.weight, br.weight, br.bias, br.running_mean, br.running_var)
        else:
            res = (None, None, None, None, None)
        if not AF.in_trunk():
            x = AF.JoinSideFn.apply(x)      # a block used on its own joins its weight-gradient stream itself
        return AF.ResBlockFn.apply(
            x, l[0].weight, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
            l[3].weight, bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var,
            *res, self.stride, training, bn1.momentum)
