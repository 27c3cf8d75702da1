"""Module-level autograd Functions of the hot path.  Each Function is one reference nn.Module.forward with a
hand-written backward; every arithmetic step is a kernel of libavec_b200.so (ops.py) - PyTorch only owns the tensors.

Reference modules mirrored (file:line in /root/reference):
  FFNFn        FeedForwardModule + the half-step residual      nnet/modules.py:257-289, nnet/blocks.py:292,301
  AttentionFn  AttentionModule + RelPos(Patch)1dMultiHeadAttention  nnet/modules.py:291-339, nnet/attentions.py:215-382
  ConvModuleFn ConvolutionModule + conv_res                     nnet/modules.py:341-385, nnet/blocks.py:273-298
  LayerNormFn  block norm                                        nnet/blocks.py:267,304
  InterCTCFn   InterCTCResModule                                 nnet/modules.py:387-400
  LinearFn / MLPFn  layers.Linear, FusionModule                  nnet/layers.py:29-76, nnet/modules.py:402-426
  AudioStemFn  AudioPreprocessing + Conv2d/BN2d/Swish stem       nnet/preprocessing.py:57-85, nnet/networks.py:359-368
  VideoStemFn  Conv3d/BN3d/ReLU + MaxPool3d                      nnet/networks.py:459-471
  ResBlockFn   ResNetBlock                                       nnet/blocks.py:29-91
  AvgPoolFn    GlobalAvgPool2d                                   nnet/networks.py:129-132
"""
import os

import torch
from torch.autograd import Function

from . import _lib as L
from . import ops
from . import weights as WL

# ------------------------------------------------------------------------------------------------------------- config
_COMPUTE_DTYPE = torch.bfloat16


def set_compute_dtype(dt):
    """torch.bfloat16 (production: tcgen05 tensor cores), torch.float32 (parity mode: exact fp32 accumulation) or "auto":
    bf16 inside a torch autocast region (the reference's Model.train_step wraps the forward in one when `precision` is a
    half type, nnet/model.py:356-360), fp32 otherwise - what avec_b200.patch_reference() selects."""
    global _COMPUTE_DTYPE
    assert dt in (torch.bfloat16, torch.float32, "auto")
    _COMPUTE_DTYPE = dt
    invalidate_weights()


_FORCED_DTYPE = None      # set while a Function.backward runs: the dtype its forward computed in


def _bwd(fn):
    """backward of a Function whose forward recorded ctx.cdt: weight copies are looked up in the FORWARD's compute dtype (with
    set_compute_dtype("auto") the ambient autocast state is gone by the time autograd runs the backward, model.py:356-367)"""
    def wrapper(ctx, *grads):
        global _FORCED_DTYPE
        prev, _FORCED_DTYPE = _FORCED_DTYPE, getattr(ctx, "cdt", None)
        try:
            return fn(ctx, *grads)
        finally:
            _FORCED_DTYPE = prev
    return wrapper


def compute_dtype():
    if _FORCED_DTYPE is not None:
        return _FORCED_DTYPE
    if _COMPUTE_DTYPE == "auto":
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32
    return _COMPUTE_DTYPE


# Compute-dtype, kernel-layout copies of the fp32 master parameters live in avec_b200.weights.PLAN: persistent buffers refreshed
# by ONE multi-tensor launch per step.  They are stale after invalidate_weights() - called by the fused optimizer (which updates
# the flat parameter buffer from a raw kernel, invisible to torch's version counters), by set_compute_dtype and by bench.py at
# the start of every timed step; parameters changed through torch are detected by version counter / storage pointer.
def invalidate_weights():
    """mark every cached compute-dtype weight copy stale (avec_b200.nnet.optimizers.Adam.step, checkpoint loads, ...)"""
    WL.PLAN.invalidate()


def new_step(arena_numel=0, device=None, advance_rng=False):
    """start of a forward pass: restart the dropout site counter, (optionally) allocate the zero-initialised fp32 arena that
    gradient accumulators and BatchNorm statistics of this step are carved from, and with advance_rng bump the device-side RNG
    step (training passes: fresh dropout / SpecAugment draws, also per CUDA-graph replay) and snapshot {seed, step} for the
    Functions of this pass (their backward regenerates the masks from the snapshot, not from the live counter)."""
    ops.RNG.site = 0
    if arena_numel and device is not None:
        ops.ARENA.begin(arena_numel, device)
    if device is not None and torch.device(device).type == "cuda":
        if advance_rng:
            ops.RNG.advance(device)
        ops.RNG.snapshot(device)


# BatchNorm.num_batches_tracked of every training-mode BatchNorm touched by a forward pass: bumped by ONE multi-tensor launch when
# the pass closes (45 one-element add kernels per AV step otherwise)
_nbt = []


def count_batch(bn):
    if _depth > 0:
        _nbt.append(bn.num_batches_tracked)
    else:
        bn.num_batches_tracked.add_(1)


# The outermost module of a forward pass (a zoo Model, or an encoder called directly / from the reference's zoo models after
# avec_b200.patch_reference()) opens the step; nested encoders see depth > 0 and do nothing.
_depth = 0


class forward_scope:
    def __init__(self, module, device):
        self.module, self.device = module, torch.device(device) if device is not None else None

    def __enter__(self):
        global _depth
        if _depth == 0:
            m = self.module
            numel = getattr(m, "_arena_numel", None)
            if numel is None:
                # parameter gradients + BatchNorm statistics / reduction scratch + positional-embedding gradients
                numel = int(1.3 * sum(p.numel() for p in m.parameters())) + (8 << 20)
                try:
                    object.__setattr__(m, "_arena_numel", numel)
                except Exception:  # noqa: BLE001
                    pass
            on_gpu = self.device is not None and self.device.type == "cuda"
            if on_gpu:
                join_side_streams(self.device)      # (only matters when a trunk input did not require a gradient)
            new_step(numel if on_gpu else 0, self.device, advance_rng=bool(m.training))
            if on_gpu and WL.PLAN.stale():
                WL.PLAN.refresh_all()      # all weight layouts in one launch, on the caller's stream, before any branch forks off
        _depth += 1
        return self

    def __exit__(self, *exc):
        global _depth
        _depth -= 1
        if _depth == 0:
            ops.RNG.snap.clear()      # Functions of this pass keep their own reference (ctx.rng)
            if _nbt:
                if exc[0] is None:
                    torch._foreach_add_(list(_nbt), 1)
                _nbt.clear()
        return False


def manual_seed(seed):
    """seed of the dropout / SpecAugment generator (Philox key); the step counter restarts at 0.  Under torch.distributed the
    rank is folded into the key (every replica draws its own masks, as torch's per-process generators do)."""
    ops.RNG.manual_seed(seed)


def wc(param, tag="plain", fn=None, pad=True):
    """compute-dtype copy of a parameter in the layout a kernel wants; fn: parameter -> VIEW of it (<= 4-d) whose row-major
    order is the [N, K] layout (default: reshape to [N, -1]).  pad: GEMM operands get the TMA-able row pitch of ops.row_pitch;
    kernels that index the weight as a dense array pass pad=False."""
    return WL.PLAN.layout((param,), (tag, pad), WL.b_plain(fn, pad), compute_dtype())


def wc_cat(params, tag):
    """parameters stacked along rows ([sum N_i, K] in the compute dtype; 1-d parameters: fp32 [sum N_i])"""
    if params[0].dim() == 1:
        return WL.PLAN.layout(tuple(params), tag, WL.b_cat(pad=False), torch.float32).reshape(-1)
    return WL.PLAN.layout(tuple(params), tag, WL.b_cat(), compute_dtype())


def wc_heads(params, tag, H, d, dp, cols=False):
    """padded-heads layout of the tcgen05 attention kernel: [H*d, K] weights (or [H*d] biases) stacked along rows with every head
    zero-padded to dp rows; cols=True: an [N, H*d] weight with every head zero-padded to dp COLUMNS (the output projection)"""
    if cols:
        return WL.PLAN.layout(tuple(params), (tag, dp), WL.b_heads_cols(H, d, dp), compute_dtype())
    if params[0].dim() == 1:
        return WL.PLAN.layout(tuple(params), (tag, dp), WL.b_heads_rows(H, d, dp, pad=False), torch.float32).reshape(-1)
    return WL.PLAN.layout(tuple(params), (tag, dp), WL.b_heads_rows(H, d, dp), compute_dtype())


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _unpad_rows(w, H, d, dp):
    """[n*H*dp, ...] fp32 gradient in the padded-heads layout -> [n*H*d, ...] (one launch; a view when dp == d)"""
    if dp == d:
        return w
    return ops.unpad_heads(w, H, d, dp, cols=False)


def _unpad_cols(w, H, d, dp):
    """[N, H*dp] -> [N, H*d]"""
    if dp == d:
        return w
    return ops.unpad_heads(w, H, d, dp, cols=True)


# --------------------------------------------------------------------------------------------------------------- FFN
class FFNFn(Function):
    """y = x + 0.5 * drop_o(W2 drop_i(swish(W1 LN(x) + b1)) + b2)   (p_in / p_out = 0: no dropout kernels are launched)"""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, w2, b2, p_in=0.0, p_out=0.0):
        ctx.cdt = compute_dtype()
        B, T, D = x.shape
        x = _c(x)
        xn, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b)
        ctx.rng = ops.RNG.cur(x.device)
        # both nn.Dropout layers ride in the epilogue of the GEMM in front of them (same Philox mask as avec_dropout)
        s_in = ops.RNG.next_site() if p_in > 0 else 0
        s_out = ops.RNG.next_site() if p_out > 0 else 0
        h, pre = ops.linear_fwd(xn.view(B * T, D), wc(w1), b1, L.EPI_SWISH, want_pre=True, drop=(ctx.rng, p_in, s_in))
        y = ops.linear_fwd(h, wc(w2), b2, L.EPI_RESIDUAL, alpha=0.5, aux=x.view(B * T, D), drop=(ctx.rng, p_out, s_out))
        ctx.save_for_backward(x, ln_w, mean, rstd, xn, pre, h, w1, w2)
        ctx.drop = (p_in, s_in, p_out, s_out)
        return y.view(B, T, D)

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, ln_w, mean, rstd, xn, pre, h, w1, w2 = ctx.saved_tensors
        p_in, s_in, p_out, s_out = ctx.drop
        B, T, D = x.shape
        dy = _c(dy)
        dy2 = dy.view(B * T, D)
        if p_out > 0:
            dyd, a = ops.dropout_rng(ctx.rng, dy2, p_out, s_out, alpha=0.5, pad_out=True), 1.0
        else:
            dyd, a = dy2, 0.5
        dpre = ops.linear_dgrad(dyd, wc(w2), L.EPI_DSWISH, alpha=a, aux=pre, drop=(ctx.rng, p_in, s_in))
        dw2 = ops.linear_wgrad(dyd, h, alpha=a)
        db2 = ops.colsum(dyd, a)
        dxn = ops.linear_dgrad(dpre, wc(w1))
        dw1 = ops.linear_wgrad(dpre, xn.view(B * T, D))
        db1 = ops.colsum(dpre)
        dx, dg, db = ops.layernorm_bwd(dxn.view(B, T, D), x, ln_w, mean, rstd, dres=dy, res_stride=1)
        return dx, dg, db, dw1, db1, dw2, db2, None, None


# --------------------------------------------------------------------------------------------------------- attention
class AttentionFn(Function):
    """y = x + upsample_P( Wo attn( pool_P(LN(x)) ) + bo )   with relative-position scores (P = 1: regular RelPos1d).
    bf16: the tcgen05 / TMEM / TMA flash kernel (csrc/attention_tc.cu) on the padded-heads layout - the Q/K/V/position
    projections write each head into its own 64- or 128-column block (zero-padded weight rows), the output projection reads
    it back through zero-padded weight columns; nothing of size T x T is stored.  fp32 parity mode / AVEC_ATTN_TC=0: the
    round-1 kernels with saved probabilities."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, wq, bq, wk, bk, wv, bv, wo, bo, wp, bp, pe, klen, H, P, p_drop=0.0):
        ctx.cdt = compute_dtype()
        B, T, D = x.shape
        d = D // H
        x = _c(x)
        ctx.rng = ops.RNG.cur(x.device)
        xp, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, P=P)
        Tp = xp.shape[1]
        dp = ops.attn_head_pad(d)
        tc = ops.ATTN_TC and x.dtype == torch.bfloat16 and dp is not None and ops.GEMM_IMPL != L.IMPL_SIMT
        if P > 1:
            klen_p = torch.div(klen, P, rounding_mode="floor").to(torch.int32) if klen is not None else None
            qlen = T // P
        else:
            klen_p, qlen = klen, Tp
        if tc:
            wqkv = wc_heads((wq, wk, wv), "qkv_tc", H, d, dp)
            bqkv = wc_heads((bq, bk, bv), "bqkv_tc", H, d, dp)
            wpp = wc_heads((wp,), "pos_tc", H, d, dp)
            bpp = wc_heads((bp,), "bpos_tc", H, d, dp)
            wop = wc_heads((wo,), "out_tc", H, d, dp, cols=True)
            qkv = ops.linear_fwd(xp.view(B * Tp, D), wqkv, bqkv)
            e = ops.linear_fwd(pe, wpp, bpp)
            o, aux = ops.relpos_attn_tc_fwd(qkv, e, klen_p, qlen, B, Tp, H, d, dp)
        else:
            wqkv = wc_cat((wq, wk, wv), "qkv")
            bqkv = wc_cat((bq, bk, bv), "bqkv")
            wop = wc(wo)
            qkv = ops.linear_fwd(xp.view(B * Tp, D), wqkv, bqkv)
            e = ops.linear_fwd(pe, wc(wp), bp)
            o, aux = ops.relpos_attn_fwd(qkv, e, klen_p, qlen, B, Tp, H, d)
        site = 0
        plain = ln_w is None          # attention.forwardQKV: no LayerNorm in front, no residual behind (modules.py:330)
        res = None if plain else x.view(B * T, D)
        if p_drop > 0 and P == 1:
            site = ops.RNG.next_site()
            y = ops.linear_fwd(o, wop, bo, L.EPI_LINEAR if plain else L.EPI_RESIDUAL, aux=res, drop=(ctx.rng, p_drop, site)).view(B, T, D)
        elif p_drop > 0 or (plain and P > 1):
            # AttentionModule.dropout acts on the upsampled (B, T, D) output: one mask element per frame (modules.py:333)
            site = ops.RNG.next_site() if p_drop > 0 else 0
            proj = ops.linear_fwd(o, wop, bo)
            y = ops.dropout_rng(ctx.rng, proj, p_drop, site, res=res, up=(T, Tp, P) if P > 1 else None).view(B, T, D)
        elif P == 1:
            y = ops.linear_fwd(o, wop, bo, L.EPI_LINEAR if plain else L.EPI_RESIDUAL, aux=res).view(B, T, D)
        else:
            proj = ops.linear_fwd(o, wop, bo)
            y = ops.upsample_add(x, proj.view(B, Tp, D), P)
        ctx.save_for_backward(x, ln_w, mean, rstd, xp, qkv, e, aux, o, pe, wq, wk, wv, wo, wp, klen_p)
        ctx.H, ctx.P, ctx.drop, ctx.tc, ctx.qlen = H, P, (p_drop, site), tc, qlen
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, ln_w, mean, rstd, xp, qkv, e, aux, o, pe, wq, wk, wv, wo, wp, klen_p = ctx.saved_tensors
        H, P, tc = ctx.H, ctx.P, ctx.tc
        B, T, D = x.shape
        Tp = xp.shape[1]
        d = D // H
        dy = _c(dy)
        p_drop, site = ctx.drop
        dyd = ops.dropout_rng(ctx.rng, dy.view(B * T, D), p_drop, site, pad_out=(P == 1)).view(B, T, D) if p_drop > 0 else dy
        dproj = dyd.view(B * T, D) if P == 1 else ops.pool_sum(dyd, P).view(B * Tp, D)
        dbo = ops.colsum(dproj)
        if tc:
            dp = ops.attn_head_pad(d)
            wqkv = wc_heads((wq, wk, wv), "qkv_tc", H, d, dp)
            wop = wc_heads((wo,), "out_tc", H, d, dp, cols=True)
            do = ops.linear_dgrad(dproj, wop)
            dwo = _unpad_cols(ops.linear_wgrad(dproj, o), H, d, dp)
            dqkv, de = ops.relpos_attn_tc_bwd(do, qkv, e, o, aux, klen_p, ctx.qlen, B, Tp, H, d, dp)
            dwp = _unpad_rows(ops.linear_wgrad(ops.convert(de, x.dtype), pe), H, d, dp)
            dbp = _unpad_rows(ops.colsum(de), H, d, dp)
            dxp = ops.linear_dgrad(dqkv, wqkv)
            dwqkv = _unpad_rows(ops.linear_wgrad(dqkv, xp.view(B * Tp, D)), H, d, dp)
            dbqkv = _unpad_rows(ops.colsum(dqkv), H, d, dp)
        else:
            do = ops.linear_dgrad(dproj, wc(wo))
            dwo = ops.linear_wgrad(dproj, o)
            dqkv, de, _, _ = ops.relpos_attn_bwd(do, qkv, e, aux, B, Tp, H, d)
            dwp = ops.linear_wgrad(ops.convert(de, x.dtype), pe)
            dbp = ops.colsum(de)
            wqkv = wc_cat((wq, wk, wv), "qkv")
            dxp = ops.linear_dgrad(dqkv, wqkv)
            dwqkv = ops.linear_wgrad(dqkv, xp.view(B * Tp, D))
            dbqkv = ops.colsum(dqkv)
        dx, dg, db = ops.layernorm_bwd(dxp.view(B, Tp, D), x, ln_w, mean, rstd, P=P, dres=None if ln_w is None else dy, res_stride=1)
        return (dx, dg, db, dwqkv[:D], dbqkv[:D], dwqkv[D:2 * D], dbqkv[D:2 * D], dwqkv[2 * D:], dbqkv[2 * D:], dwo, dbo,
                dwp, dbp, None, None, None, None, None)


class GroupedAttentionFn(Function):
    """y = x + Wo attn_G(LN(x)) + bo: Transformer-XL style relative attention with content / position biases u, v over
    tokens made of G consecutive frames (GroupedRelPosMultiHeadSelfAttention, reference nnet/attentions.py:579-650;
    G = 1 is RelPosMultiHeadSelfAttention).  Projections run at full frame rate.  bf16: the frame-rate q / k / v are regrouped
    into the padded-heads token layout (4 parts: q + u | k | v | q + v, one kernel) and the tcgen05 flash kernel of
    csrc/attention_tc.cu does the rest; fp32 parity mode: the round-1 SIMT kernels, grouping as pure addressing."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, wq, bq, wk, bk, wv, bv, wo, bo, wp, bp, u, v, pe, klen, H, G, p_drop=0.0):
        ctx.cdt = compute_dtype()
        B, T, D = x.shape
        Tn = -(-T // G)
        d = G * D // H
        x = _c(x)
        ctx.rng = ops.RNG.cur(x.device)
        xn, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b)
        wqkv = wc_cat((wq, wk, wv), "qkv")
        bqkv = wc_cat((bq, bk, bv), "bqkv")
        qkv = ops.linear_fwd(xn.view(B * T, D), wqkv, bqkv)
        e = ops.linear_fwd(pe, wc(wp), bp)                      # [2*Tp-G, D] == [2*Tn-1, G*D]
        klen_g = torch.div(klen + (G - 1), G, rounding_mode="floor").to(torch.int32) if klen is not None else None
        dp = ops.attn_head_pad(d)
        tc = ops.ATTN_TC and x.dtype == torch.bfloat16 and dp is not None and ops.GEMM_IMPL != L.IMPL_SIMT
        if tc:
            tok = ops.attn_group_pack(qkv, u, v, B, T, Tn, G, H, D, dp, 4)
            e_tok = ops.attn_group_pack(e, None, None, 1, e.shape[0], 2 * Tn - 1, G, H, D, dp, 1)
            o_tok, aux = ops.relpos_attn_tc_fwd(tok, e_tok, klen_g, Tn, B, Tn, H, d, dp, qp_part=3)
            o = ops.attn_group_unpack(o_tok, B, T, Tn, G, H, D, dp)
            saved_qkv, saved_e = tok, e_tok
        else:
            o, aux = ops.relpos_attn_fwd(qkv, e, klen_g, Tn, B, Tn, H, d, G=G, Tf=T, u=u, v=v)
            o_tok, saved_qkv, saved_e = None, qkv, e
        site = 0
        plain = ln_w is None
        if p_drop > 0:
            site = ops.RNG.next_site()
            y = ops.linear_fwd(o, wc(wo), bo, L.EPI_LINEAR if plain else L.EPI_RESIDUAL, aux=None if plain else x.view(B * T, D),
                               drop=(ctx.rng, p_drop, site)).view(B, T, D)
        elif plain:
            y = ops.linear_fwd(o, wc(wo), bo).view(B, T, D)
        else:
            y = ops.linear_fwd(o, wc(wo), bo, L.EPI_RESIDUAL, aux=x.view(B * T, D)).view(B, T, D)
        ctx.save_for_backward(x, ln_w, mean, rstd, xn, saved_qkv, saved_e, aux, o, o_tok, pe, wq, wk, wv, wo, wp, u, v, klen_g)
        ctx.H, ctx.G, ctx.drop, ctx.tc = H, G, (p_drop, site), tc
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, ln_w, mean, rstd, xn, qkv, e, aux, o, o_tok, pe, wq, wk, wv, wo, wp, u, v, klen_g = ctx.saved_tensors
        H, G = ctx.H, ctx.G
        B, T, D = x.shape
        Tn = -(-T // G)
        d = G * D // H
        dy = _c(dy)
        p_drop, site = ctx.drop
        dy2 = ops.dropout_rng(ctx.rng, dy.view(B * T, D), p_drop, site, pad_out=True) if p_drop > 0 else dy.view(B * T, D)
        do = ops.linear_dgrad(dy2, wc(wo))
        dwo = ops.linear_wgrad(dy2, o)
        dbo = ops.colsum(dy2)
        if ctx.tc:
            dp = ops.attn_head_pad(d)
            do_tok = ops.attn_group_pack(do, None, None, B, T, Tn, G, H, D, dp, 1)
            dqkv_tok, de_tok = ops.relpos_attn_tc_bwd(do_tok, qkv, e, o_tok, aux, klen_g, Tn, B, Tn, H, d, dp, qp_part=3)
            dqkv, du, dv = ops.attn_group_unpack_dqkv(dqkv_tok, B, T, Tn, G, H, D, dp)
            de2 = ops.attn_group_unpack(de_tok, 1, pe.shape[0], 2 * Tn - 1, G, H, D, dp)
        else:
            dqkv, de, du, dv = ops.relpos_attn_bwd(do, qkv, e, aux, B, Tn, H, d, G=G, Tf=T, u=u, v=v)
            de2 = de.view(-1, D)
        dwp = ops.linear_wgrad(ops.convert(de2, x.dtype), pe)
        dbp = ops.colsum(de2)
        wqkv = wc_cat((wq, wk, wv), "qkv")
        dxn = ops.linear_dgrad(dqkv, wqkv)
        dwqkv = ops.linear_wgrad(dqkv, xn.view(B * T, D))
        dbqkv = ops.colsum(dqkv)
        dx, dg, db = ops.layernorm_bwd(dxn.view(B, T, D), x, ln_w, mean, rstd, dres=None if ln_w is None else dy, res_stride=1)
        return (dx, dg, db, dwqkv[:D], dbqkv[:D], dwqkv[D:2 * D], dbqkv[D:2 * D], dwqkv[2 * D:], dbqkv[2 * D:], dwo, dbo,
                dwp, dbp, du, dv, None, None, None, None, None)


# ------------------------------------------------------------------------------------------------------- conv module
class ConvModuleFn(Function):
    """y = res(x) + W3 swish(BN(dwconv_k15,s(GLU(W1 LN(x) + b1)))) + b3,  res = identity | Conv1d(k=1, stride s)"""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, wd, bd, bn_w, bn_b, rm, rv, w3, b3, wr, br, stride, training, momentum, p_drop=0.0):
        ctx.cdt = compute_dtype()
        B, T, D = x.shape
        De = w3.shape[0]
        ks = wd.shape[-1]
        x = _c(x)
        ctx.rng = ops.RNG.cur(x.device)
        xn, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b)
        pre = ops.linear_fwd(xn.view(B * T, D), wc(w1), b1).view(B, T, 2 * De)
        wdw = wd.detach().reshape(De, ks)
        u, stats = ops.glu_dwconv_fwd(pre, wdw, bd, stride, ks, want_stats=training)
        To = u.shape[1]
        u2 = u.view(B * To, De)
        if training:
            bnbuf = ops.bn_finalize(stats, bn_w, bn_b, B * To, rm, rv, 1e-5, momentum)
        else:
            bnbuf = ops.bn_eval_affine(bn_w, bn_b, rm, rv, 1e-5)
        v = ops.bn_apply(u2, bnbuf[0], bnbuf[1], L.ACT_SWISH, pad_out=True)     # operand of the pointwise conv and of its wgrad
        if wr is None:
            xs = None
            aux = x.view(B * T, D)
        else:
            xs = _c(x[:, ::stride]).view(B * To, D) if stride > 1 else x.view(B * T, D)
            aux = ops.linear_fwd(xs, wc(wr), br)
        site = 0
        if p_drop > 0:
            site = ops.RNG.next_site()
            y = ops.linear_fwd(v, wc(w3), b3, L.EPI_RESIDUAL, aux=aux, drop=(ctx.rng, p_drop, site)).view(B, To, De)
        else:
            y = ops.linear_fwd(v, wc(w3), b3, L.EPI_RESIDUAL, aux=aux).view(B, To, De)
        ctx.save_for_backward(x, ln_w, mean, rstd, xn, pre, u, bnbuf, v, xs, w1, wd, bn_w, w3, wr)
        ctx.stride, ctx.training, ctx.drop = stride, training, (p_drop, site)
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, ln_w, mean, rstd, xn, pre, u, bnbuf, v, xs, w1, wd, bn_w, w3, wr = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("avec_b200: ConvModule backward needs training-mode BatchNorm statistics")
        stride = ctx.stride
        B, T, D = x.shape
        To, De = u.shape[1], u.shape[2]
        ks = wd.shape[-1]
        dy = _c(dy)
        dy2 = dy.view(B * To, De)
        p_drop, site = ctx.drop
        dyd = ops.dropout_rng(ctx.rng, dy2, p_drop, site, pad_out=True) if p_drop > 0 else dy2
        dv = ops.linear_dgrad(dyd, wc(w3))
        dw3 = ops.linear_wgrad(dyd, v)
        db3 = ops.colsum(dyd)
        du, _, dgamma, dbeta = ops.bn_bwd(dv, u.view(B * To, De), bnbuf, bn_w, L.ACT_SWISH)
        dpre, dwd, dbd = ops.glu_dwconv_bwd(du.view(B, To, De), pre, wd.detach().reshape(De, ks), stride, ks)
        dpre2 = dpre.view(B * T, 2 * De)
        dxn = ops.linear_dgrad(dpre2, wc(w1))
        dw1 = ops.linear_wgrad(dpre2, xn.view(B * T, D))
        db1 = ops.colsum(dpre2)
        if wr is None:
            dres, dwr, dbr = dy, None, None
        else:
            dres = ops.linear_dgrad(dy2, wc(wr)).view(B, To, D)
            dwr = ops.linear_wgrad(dy2, xs).view(wr.shape)
            dbr = db3.clone() if p_drop == 0 else ops.colsum(dy2)
        dx, dg, db = ops.layernorm_bwd(dxn.view(B, T, D), x, ln_w, mean, rstd, dres=dres, res_stride=stride)
        return (dx, dg, db, dw1.view(w1.shape), db1, dwd.view(wd.shape), dbd, dgamma, dbeta, None, None, dw3.view(w3.shape), db3,
                dwr, dbr, None, None, None, None)


class DropoutFn(Function):
    """nn.Dropout on a (.., C) tensor (ConformerInterCTC input dropout, networks.py:269)"""

    @staticmethod
    def forward(ctx, x, p):
        ctx.cdt = compute_dtype()
        site = ops.RNG.next_site()
        ctx.drop = (p, site)
        ctx.rng = ops.RNG.cur(x.device)
        return ops.dropout_rng(ctx.rng, _c(x), p, site)

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        p, site = ctx.drop
        return ops.dropout_rng(ctx.rng, _c(dy), p, site), None


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.cdt = compute_dtype()
        x = _c(x)
        y, mean, rstd = ops.layernorm_fwd(x, w, b, pad_out=False)     # the block output: read by row kernels, dense rows
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(_c(dy), x, w, mean, rstd)
        return dx, dg, db


# ----------------------------------------------------------------------------------------------------------- InterCTC
class InterCTCFn(Function):
    """logits = W1 x + b1 (fp32);  y = x + W2 softmax(logits) + b2"""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        ctx.cdt = compute_dtype()
        B, T, D = x.shape
        x = _c(x)
        x2 = x.view(B * T, D)
        logits = ops.linear_fwd(x2, wc(w1), b1, out_dtype=torch.float32)
        p = ops.softmax_fwd(logits, x.dtype)
        y = ops.linear_fwd(p, wc(w2), b2, L.EPI_RESIDUAL, aux=x2)
        ctx.save_for_backward(x, p, w1, w2)
        return y.view(B, T, D), logits.view(B, T, -1)

    @staticmethod
    @_bwd
    def backward(ctx, dy, dlogits):
        x, p, w1, w2 = ctx.saved_tensors
        B, T, D = x.shape
        x2 = x.view(B * T, D)
        dy2 = _c(dy).view(B * T, D)
        dp = ops.linear_dgrad(dy2, wc(w2))
        dw2 = ops.linear_wgrad(dy2, p)
        db2 = ops.colsum(dy2)
        dadd = _c(dlogits.float()).view(B * T, -1) if dlogits is not None else None
        dl = ops.softmax_bwd(dp, p, dadd, out_dtype=x.dtype)
        dx = ops.linear_dgrad(dl, wc(w1), L.EPI_RESIDUAL, aux=dy2)
        dw1 = ops.linear_wgrad(dl, x2)
        db1 = ops.colsum(dl)
        return dx.view(B, T, D), dw1, db1, dw2, db2


# ------------------------------------------------------------------------------------------------------------- Linear
class LinearFn(Function):
    """y = x W^T + b on the last dim.  wl = optional (tag, to_kernel_layout, grad_to_param_layout) for permuted weights."""

    @staticmethod
    def forward(ctx, x, w, b, out_fp32, wl):
        ctx.cdt = compute_dtype()
        shp = x.shape
        x2 = _c(x).view(-1, shp[-1])
        wk = wc(w, wl[0], wl[1]) if wl is not None else wc(w)
        y = ops.linear_fwd(x2, wk, b, out_dtype=torch.float32 if out_fp32 else None)
        ctx.save_for_backward(x2, w)
        ctx.shp, ctx.wl, ctx.has_b = shp, wl, b is not None
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        wl = ctx.wl
        wk = wc(w, wl[0], wl[1]) if wl is not None else wc(w)
        dy2 = _c(dy).view(-1, w.shape[0])
        if dy2.dtype != x2.dtype:
            dy2 = ops.convert(dy2, x2.dtype)
        dx = ops.linear_dgrad(dy2, wk).view(ctx.shp)
        dw = ops.linear_wgrad(dy2, x2)
        if wl is not None:
            dw = wl[2](dw)
        db = ops.colsum(dy2) if ctx.has_b else None
        return dx, dw.view(w.shape), db, None, None


class MLPFn(Function):
    """y = W2 swish(W1 x + b1) + b2   (FusionModule body)"""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        ctx.cdt = compute_dtype()
        shp = x.shape
        x2 = _c(x).view(-1, shp[-1])
        h, pre = ops.linear_fwd(x2, wc(w1), b1, L.EPI_SWISH, want_pre=True)
        y = ops.linear_fwd(h, wc(w2), b2)
        ctx.save_for_backward(x2, pre, h, w1, w2)
        ctx.shp = shp
        return y.view(*shp[:-1], w2.shape[0])

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x2, pre, h, w1, w2 = ctx.saved_tensors
        dy2 = _c(dy).view(-1, w2.shape[0])
        dpre = ops.linear_dgrad(dy2, wc(w2), L.EPI_DSWISH, alpha=1.0, aux=pre)
        dw2 = ops.linear_wgrad(dy2, h)
        db2 = ops.colsum(dy2)
        dx = ops.linear_dgrad(dpre, wc(w1)).view(ctx.shp)
        dw1 = ops.linear_wgrad(dpre, x2)
        db1 = ops.colsum(dpre)
        return dx, dw1, db1, dw2, db2


# --------------------------------------------------------------------------------------------------------- front-ends
def _bn_buf(stats, w, b, rm, rv, count, training, momentum):
    if training:
        return ops.bn_finalize(stats, w, b, count, rm, rv, 1e-5, momentum)
    return ops.bn_eval_affine(w, b, rm, rv, 1e-5)


class AudioStemFn(Function):
    """wave [B,L] -> log-mel [B,F,80] -> Conv2d(1->C, k3, s2, same) + BN2d + Swish -> [B, F', 40*C] (feature = f*C + c)"""

    @staticmethod
    def forward(ctx, wave, fb, cw, cb, bn_w, bn_b, rm, rv, training, momentum, spec=None, mel_len=None):
        ctx.cdt = compute_dtype()
        B = wave.shape[0]
        mel = ops.stft_mel_log(_c(wave.float()), fb, layout=0)
        if spec is not None:      # SpecAugment (mF, F, mT, pS) on the fp32 log-mel, in place (networks.py:423-424)
            ops.spec_augment_(mel, mel_len, ops.RNG.next_site(), *spec, rng=ops.RNG.cur(mel.device))
        F = mel.shape[1]
        melc = ops.convert(mel, compute_dtype())
        Co = cw.shape[0]
        g = ops.make_geom(B, 1, F, 80, 1, Co, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        wp = wc(cw, "stem2d", lambda w: w[:, 0].transpose(1, 2), pad=False)
        sites = ops.geom_sites(g)
        stats = ops.gemm_stats_buffer(Co, wave.device) if training else None
        direct = Co % 4 == 0 and Co <= 256   # SIMT stem kernel (K = 9 is far below a tensor-core tile; output-write bound)
        u = ops.stem2d_fwd(melc, wp, cb, colstats=stats) if direct else ops.conv_fwd(melc, wp, g, bias=cb, colstats=stats)
        bnbuf = _bn_buf(stats, bn_w, bn_b, rm, rv, sites, training, momentum)
        v = ops.bn_apply(u, bnbuf[0], bnbuf[1], L.ACT_SWISH)
        ctx.save_for_backward(melc, u, bnbuf, bn_w, cw)
        ctx.g, ctx.training, ctx.direct = g, training, direct
        return v.view(B, g.Ho, g.Wo * Co)

    @staticmethod
    @_bwd
    def backward(ctx, dv):
        melc, u, bnbuf, bn_w, cw = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("avec_b200: stem backward needs training-mode BatchNorm statistics")
        g = ctx.g
        Co = cw.shape[0]
        dv2 = _c(dv).view(-1, Co)
        du, _, dgamma, dbeta = ops.bn_bwd(dv2, u, bnbuf, bn_w, L.ACT_SWISH)
        dwp = ops.stem2d_wgrad(melc, du) if ctx.direct else ops.conv_wgrad(du, melc, g)
        dcw = dwp.view(Co, 3, 3).transpose(1, 2).reshape(cw.shape)
        dcb = ops.colsum(du)
        return None, None, dcw, dcb, dgamma, dbeta, None, None, None, None, None, None


class VideoStemFn(Function):
    """video [B,T,H,W,1] -> Conv3d(1->64,(5,7,7),s(1,2,2),same)+BN3d+ReLU -> MaxPool3d((1,3,3),s(1,2,2),same) -> [B*T,H/4,W/4,64]"""

    @staticmethod
    def forward(ctx, video, cw, cb, bn_w, bn_b, rm, rv, training, momentum):
        ctx.cdt = compute_dtype()
        B, T, H, W = video.shape[0], video.shape[1], video.shape[2], video.shape[3]
        xc = ops.convert(_c(video.float()).view(B * T * H, W), compute_dtype()).view(B, T, H, W, 1)
        Co = cw.shape[0]
        kt, kh, kw = cw.shape[2], cw.shape[3], cw.shape[4]
        g = ops.make_geom(B, T, H, W, 1, Co, (kt, kh, kw), (1, 2, 2), ((kt - 1) // 2, (kh - 1) // 2, (kw - 1) // 2))
        taps = kt * kh * kw
        sites = ops.geom_sites(g)
        stats = ops.gemm_stats_buffer(Co, video.device) if training else None
        direct = ops.stem3d_supported(xc, Co, kt, kh, kw)
        if direct:
            # direct tcgen05 implicit GEMM: the im2col tile only ever exists in shared memory
            u = ops.stem3d_fwd(xc, ops.stem3d_weight_layout(cw), cb, colstats=stats)
            col = xc
        else:
            # fallback (fp32 parity mode / other geometries): im2col + plain GEMMs (fwd and wgrad share the [sites, Kpad] matrix)
            Kpad = (taps + 63) // 64 * 64
            wp = WL.PLAN.layout((cw,), ("stem3d", Kpad), WL.b_custom(lambda w: w.reshape(w.shape[0], -1), Co, Kpad, (Kpad, 1)), compute_dtype())
            col = ops.im2col_c1(xc, g, Kpad)
            u = ops.linear_fwd(col, wp, cb, colstats=stats)
        bnbuf = _bn_buf(stats, bn_w, bn_b, rm, rv, sites, training, momentum)
        y, idx = ops.bn_relu_maxpool_fwd(u, bnbuf[0], bnbuf[1], B * T, g.Ho, g.Wo, Co)
        ctx.save_for_backward(col, u, bnbuf, bn_w, cw, idx)
        ctx.g, ctx.training, ctx.direct = g, training, direct
        return y

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        col, u, bnbuf, bn_w, cw, idx = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("avec_b200: stem backward needs training-mode BatchNorm statistics")
        g = ctx.g
        Co = cw.shape[0]
        # (ops.bn_bwd_pool fuses the max-pool backward into both BatchNorm passes and never materialises the 1.6 GB pre-pool
        # gradient, but its gathers make the two passes 4.25 ms against 3.2 ms for the three streaming kernels below)
        dz = ops.bn_relu_maxpool_bwd(_c(dy), idx, g.N * g.To, g.Ho, g.Wo, Co)
        du, _, dgamma, dbeta = ops.bn_bwd(dz, u, bnbuf, bn_w, L.ACT_NONE)
        taps = cw.shape[2] * cw.shape[3] * cw.shape[4]
        if ctx.direct:
            dcw = ops.stem3d_wgrad(col, du).reshape(cw.shape)
        else:
            dcw = ops.linear_wgrad(du, col)[:, :taps].reshape(cw.shape)
        dcb = ops.colsum(du)
        return None, dcw, dcb, dgamma, dbeta, None, None, None, None


def _pack_fwd(w):   # (Co, Ci, kh, kw) -> view (Co, kh, kw, Ci) = rows of the [Co, taps*Ci] forward operand
    return w.permute(0, 2, 3, 1)


def _pack_dgrad(w):  # (Co, Ci, kh, kw) -> view (Ci, kh, kw, Co) = rows of the [Ci, taps*Co] dgrad operand
    return w.permute(1, 2, 3, 0)


def _unpack_wgrad(dw, w):  # [Co, taps*Ci] -> (Co, Ci, kh, kw)
    Co, Ci, kh, kw = w.shape
    return dw.view(Co, kh, kw, Ci).permute(0, 3, 1, 2).contiguous()


class ResBlockFn(Function):
    """ResNet BasicBlock on channels-last images: relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut(x))"""

    @staticmethod
    def forward(ctx, x, w1, g1w, g1b, rm1, rv1, w2, g2w, g2b, rm2, rv2, wr, grw, grb, rmr, rvr, stride, training, momentum):
        ctx.cdt = compute_dtype()
        N, H, W, Ci = x.shape
        Co = w1.shape[0]
        x = _c(x)
        dev = x.device

        def st():
            return ops.gemm_stats_buffer(Co, dev) if training else None

        ga = ops.make_geom(N, 1, H, W, Ci, Co, (1, 3, 3), (1, stride, stride), (0, 1, 1))
        s1 = st()
        u1 = ops.conv_fwd(x, wc(w1, "cf", _pack_fwd), ga, colstats=s1)
        sites = ops.geom_sites(ga)
        bn1 = _bn_buf(s1, g1w, g1b, rm1, rv1, sites, training, momentum)
        a1 = ops.bn_apply(u1, bn1[0], bn1[1], L.ACT_RELU)
        gb = ops.make_geom(N, 1, ga.Ho, ga.Wo, Co, Co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        s2 = st()
        u2 = ops.conv_fwd(a1, wc(w2, "cf", _pack_fwd), gb, colstats=s2)
        bn2 = _bn_buf(s2, g2w, g2b, rm2, rv2, sites, training, momentum)
        if wr is not None:
            gr = ops.make_geom(N, 1, H, W, Ci, Co, (1, 1, 1), (1, stride, stride), (0, 0, 0))
            sr = st()
            ur = ops.conv_fwd(x, wc(wr, "cf", _pack_fwd), gr, colstats=sr)
            bnr = _bn_buf(sr, grw, grb, rmr, rvr, sites, training, momentum)
            r = ops.bn_apply(ur, bnr[0], bnr[1], L.ACT_NONE)
        else:
            gr, ur, bnr = None, None, None
            r = x.view(sites, Co)
        y = ops.bn_apply(u2, bn2[0], bn2[1], L.ACT_RELU, res=r)
        ctx.save_for_backward(x, u1, bn1, a1, u2, bn2, ur, bnr, r, w1, g1w, w2, g2w, wr, grw)
        ctx.geoms, ctx.training, ctx.wg_side = (ga, gb, gr), training, _wg_scoped_off == 0
        return y.view(N, ga.Ho, ga.Wo, Co)

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        x, u1, bn1, a1, u2, bn2, ur, bnr, r, w1, g1w, w2, g2w, wr, grw = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("avec_b200: ResNet backward needs training-mode BatchNorm statistics")
        ga, gb, gr = ctx.geoms
        Co = w1.shape[0]
        dy2 = _c(dy).view(-1, Co)
        # Weight gradients are off the critical path (nothing in the backward waits for them): they go to a side stream, enqueued
        # AFTER the input-gradient convolution that shares their operand, so that they run under the BatchNorm backward passes
        # that follow (tensor-pipe work under HBM-bound work; JoinSideFn at the trunk input joins the stream again).
        du2, dres, dg2, db2 = ops.bn_bwd(dy2, u2, bn2, g2w, L.ACT_RELU, res=r, want_dres=True)
        da1 = ops.conv_dgrad(du2, wc(w2, "cd", _pack_dgrad), gb)
        dw2 = _wgrad_side(du2, a1, gb, w2, ctx.wg_side)
        du1, _, dg1, db1 = ops.bn_bwd(da1, u1, bn1, g1w, L.ACT_RELU)
        if wr is not None:
            dur, _, dgr, dbr = ops.bn_bwd(dres, ur, bnr, grw, L.ACT_NONE)
            dxr = ops.conv_dgrad(dur, wc(wr, "cd", _pack_dgrad), gr)
            dx = ops.conv_dgrad(du1, wc(w1, "cd", _pack_dgrad), ga, epi=L.EPI_RESIDUAL, aux=dxr)
            dwr = _wgrad_side(dur, x, gr, wr, ctx.wg_side)
        else:
            dwr = dgr = dbr = None
            dx = ops.conv_dgrad(du1, wc(w1, "cd", _pack_dgrad), ga, epi=L.EPI_RESIDUAL, aux=dres)
        dw1 = _wgrad_side(du1, x, ga, w1, ctx.wg_side)
        return (dx.view(x.shape), dw1, dg1, db1, None, None, dw2, dg2, db2, None, None, dwr, dgr, dbr, None, None, None, None, None)


# ---- weight gradients of the ResNet trunk on a side stream -----------------------------------------------------------------
WGRAD_OVERLAP = os.environ.get("AVEC_WGRAD_OVERLAP", "1") != "0"
_wg_streams = {}
_wg_scoped_off = 0


class no_wgrad_overlap:
    """forward-time scope: the ResNet blocks built inside keep their weight gradients on the main stream.  The AV encoder uses it
    when its audio branch already runs concurrently with the video branch (measured: the third stream then costs 0.4 ms per
    step instead of saving 1.3 ms as it does for the visual-only model)"""
    def __enter__(self):
        global _wg_scoped_off
        _wg_scoped_off += 1

    def __exit__(self, *exc):
        global _wg_scoped_off
        _wg_scoped_off -= 1
        return False


def _wg_stream(device):
    s = _wg_streams.get(device)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _wg_streams[device] = s
    return s


_wg_pending = {}      # device -> tensors the side stream may still be reading (kept alive until the join)


def _wgrad_side(dyt, xt, g, w, enabled=True):
    """dW of one convolution, (Co, Ci, kh, kw) fp32, computed on the weight-gradient stream after everything enqueued so far.
    The operands stay referenced until join_side_streams(): no record_stream (with 400 MB operands it makes the caching
    allocator hold every freed block back and the reserved pool balloons)."""
    if not (enabled and WGRAD_OVERLAP and dyt.is_cuda):
        return _unpack_wgrad(ops.conv_wgrad(dyt, xt, g), w)
    dev = dyt.device
    cur = torch.cuda.current_stream(dev)
    side = _wg_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        dw = _unpack_wgrad(ops.conv_wgrad(dyt, xt, g), w)
    _wg_pending.setdefault(dev, []).extend((dyt, xt))
    return dw


def join_side_streams(device=None):
    """make the current stream wait for the weight-gradient stream (end of a backward pass); a no-op when nothing is pending"""
    for dev in list(_wg_pending.keys()):
        if device is not None and torch.device(device) != dev:
            continue
        if _wg_pending[dev]:
            torch.cuda.current_stream(dev).wait_stream(_wg_streams[dev])
            _wg_pending[dev] = []


_trunk_depth = 0     # > 0 while a ResNet trunk (which joins once, at its input) is running its blocks


class trunk_scope:
    def __enter__(self):
        global _trunk_depth
        _trunk_depth += 1

    def __exit__(self, *exc):
        global _trunk_depth
        _trunk_depth -= 1
        return False


def in_trunk():
    return _trunk_depth > 0


class JoinSideFn(Function):
    """identity placed at the input of the ResNet trunk: its backward runs after every block's backward and joins the
    weight-gradient stream back onto the stream autograd runs the trunk on"""

    @staticmethod
    def forward(ctx, x):
        ctx.cdt = compute_dtype()
        return x.view_as(x)

    @staticmethod
    @_bwd
    def backward(ctx, g):
        if g.is_cuda:
            join_side_streams(g.device)
        return g


class AvgPoolFn(Function):
    """[N, H, W, C] -> [N, C]"""

    @staticmethod
    def forward(ctx, x):
        ctx.cdt = compute_dtype()
        N, H, W, Cn = x.shape
        ctx.shp = x.shape
        return ops.avgpool_fwd(_c(x), N, H * W, Cn)

    @staticmethod
    @_bwd
    def backward(ctx, dy):
        N, H, W, Cn = ctx.shp
        return ops.avgpool_bwd(_c(dy), N, H * W, Cn).view(ctx.shp)


class CTCFn(Function):
    """per-utterance CTC negative log-likelihood of fp32 logits [B,T,V] (log-softmax fused)"""

    @staticmethod
    def forward(ctx, logits, labels, in_len, lab_len, blank, zero_infinity):
        ctx.cdt = compute_dtype()
        lg = _c(logits.float())
        dev = lg.device
        lab = _c(labels.to(device=dev, dtype=torch.long))
        il = _c(in_len.to(device=dev, dtype=torch.long)) if in_len is not None else None
        ll = _c(lab_len.to(device=dev, dtype=torch.long))
        nll, grad = ops.ctc_loss(lg, lab, il, ll, blank, zero_infinity)
        ctx.save_for_backward(grad)
        return nll

    @staticmethod
    @_bwd
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout.view(-1, 1, 1), None, None, None, None, None
