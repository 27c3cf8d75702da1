"""Tensor-level wrappers over the C ABI: they take torch CUDA tensors, allocate outputs with torch (device memory and
streams are the only things PyTorch provides here) and launch the sm_100a kernels on the current stream."""
import os

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}

# global knob: which GEMM engine avec_gemm uses (AUTO = tcgen05 whenever the operands are bf16 and the shape qualifies)
GEMM_IMPL = L.IMPL_AUTO


def set_gemm_impl(name):
    global GEMM_IMPL
    GEMM_IMPL = {"auto": L.IMPL_AUTO, "simt": L.IMPL_SIMT, "tcgen05": L.IMPL_TCGEN05}[name]


def set_tma(enabled):
    """diagnostics: False forces the gather producers of the tcgen05 GEMM even where TMA descriptors are possible"""
    L.load().avec_set_tma(1 if enabled else 0)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _dt(t):
    return _DT[t.dtype]


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("avec_b200 ops need CUDA tensors: the hot path has no CPU fallback")


# ---- zero-initialised fp32 scratch (gradient accumulators, BatchNorm statistics): one memset per step instead of ~900
class _Arena:
    def __init__(self):
        self.buf, self.off = None, 0

    def begin(self, numel, device):
        self.buf = torch.zeros(int(numel), device=device, dtype=torch.float32) if numel else None
        self.off = 0

    def take(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        if self.buf is not None and self.buf.device == torch.device(device) and self.off + n <= self.buf.numel():
            out = self.buf[self.off:self.off + n].view(*shape)
            self.off += (n + 3) // 4 * 4
            return out
        return torch.zeros(shape, device=device, dtype=torch.float32)


ARENA = _Arena()


def zeros_f32(shape, device):
    return ARENA.take(tuple(shape) if isinstance(shape, (tuple, list)) else (shape,), device)


def launch_count():
    return L.load().avec_launch_count()


def set_pdl(enabled):
    """programmatic dependent launch of the tcgen05 GEMM kernel (on by default)"""
    L.load().avec_set_pdl(int(bool(enabled)))


def pdl_exclude_stream(stream, enabled=True):
    """GEMMs launched on `stream` (torch.cuda.Stream) are serialised the ordinary way: for a secondary stream whose early-scheduled
    CTAs would take shared memory from the critical-path stream"""
    L.load().avec_pdl_exclude_stream(stream.cuda_stream if stream is not None else None, int(bool(enabled)))


def reset_launch_count():
    L.load().avec_reset_launch_count()


# --------------------------------------------------------------------------------------------------------------- GEMM
def _gemm(args, may_decline=False):
    """may_decline: AVEC_ERR_UNSUPPORTED is an answer (False), not an error - nothing was launched"""
    rc = L.load().avec_gemm(args, _stream())
    if may_decline and rc == ERR_UNSUPPORTED:
        return False
    L.check(rc, "avec_gemm")
    return True


ERR_UNSUPPORTED = -3
# nn.Dropout inside the GEMM epilogue - 0: never, 1: wherever the kernel can, 2 (default): only behind GEMMs whose output is not wider
# than twice their contraction (FFN second Linear, attention output projection, ConvModule's last pointwise conv).  The wide, short
# ones (FFN first Linear: 4D outputs from D inputs) are bound by their 4-8 epilogue warps, where the Philox draws cost more than the
# separate full-occupancy dropout kernel they replace (profiles/r02_launches_train_graph.md).
FUSE_DROPOUT = int(os.environ.get("AVEC_FUSE_DROPOUT", "2"))


def _gemm_drop(args, drop):
    """launch with nn.Dropout fused into the epilogue (drop = (rng, p, site)); False: this operand / epilogue combination has no
    fused form (fp32 parity mode, SIMT implementation, unaligned rows) and the caller runs GEMM + avec_dropout instead."""
    if not FUSE_DROPOUT or (int(FUSE_DROPOUT) == 2 and args.N > 2 * args.K):
        return False
    rng, p, site = drop
    args.drop_p, args.drop_site, args.drop_rng = float(p), int(site), rng.data_ptr()
    if not _gemm(args, True):
        args.drop_p, args.drop_rng = 0.0, None
        return False
    return True


def _epi(args, epi, alpha, bias, out, out2=None, aux=None, colstats=None):
    args.epi = epi
    args.alpha = alpha
    args.bias = _p(bias)
    args.out = out.data_ptr()
    args.out_dtype = _dt(out)
    args.ldo = out.stride(0) if out.dim() == 2 else out.shape[-1]
    if out2 is not None:
        args.out2, args.out2_dtype, args.ldo2 = out2.data_ptr(), _dt(out2), out2.stride(0)
    if aux is not None:
        args.aux, args.aux_dtype, args.ldaux = aux.data_ptr(), _dt(aux), aux.stride(0)
    args.colstats = _p(colstats)


def _split_for(M, N, K):
    """split-K factor of a weight-gradient GEMM: about one wave of CTAs, at least 8 k-blocks (512 reduction rows) each"""
    tiles = -(-M // 128) * -(-N // 192)
    kb = -(-K // 64)
    return max(1, min(kb // 8 if kb >= 16 else 1, max(1, 160 // max(1, tiles)), 512))


def linear_fwd(x, w, bias=None, epi=L.EPI_LINEAR, alpha=1.0, aux=None, out_dtype=None, want_pre=False, colstats=None, drop=None):
    """out[M,N] = epi(x[M,K] @ w[N,K]^T + bias).  x, w same dtype, row-major (w may have a padded leading dim).
    drop = (rng, p, site): nn.Dropout behind the GEMM with the mask avec_dropout draws for (rng, site) -
    LINEAR: drop(acc), SWISH: drop(swish(acc)) (the pre-activation copy stays whole), RESIDUAL: aux + alpha * drop(acc)."""
    _cuda(x, w)
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=x.device, dtype=out_dtype or x.dtype)
    pre = torch.empty((M, N), device=x.device, dtype=x.dtype) if want_pre else None
    a = L.GemmArgs()
    a.mode, a.impl, a.M, a.N, a.K = L.GEMM_PLAIN, GEMM_IMPL, M, N, K
    a.A, a.sam, a.sak = x.data_ptr(), x.stride(0), 1
    a.B, a.sbn, a.sbk = w.data_ptr(), w.stride(0), 1
    a.ab_dtype = _dt(x)
    if drop is not None and drop[1] > 0:
        _epi(a, epi, alpha, bias, out, pre, aux, colstats)
        if not _gemm_drop(a, drop):
            if epi == L.EPI_RESIDUAL:      # aux + alpha * drop(acc): plain GEMM, then the dropout kernel adds the residual
                _epi(a, L.EPI_LINEAR, 1.0, bias, out, pre, None, colstats)
                a.aux = None
                _gemm(a)
                assert aux.is_contiguous() and aux.dtype == out.dtype
                dropout_rng(drop[0], out, drop[1], drop[2], res=aux, alpha=alpha, out=out)
            else:
                assert epi in (L.EPI_LINEAR, L.EPI_SWISH)
                _gemm(a)
                dropout_rng(drop[0], out, drop[1], drop[2], out=out)
        return (out, pre) if want_pre else out
    _epi(a, epi, alpha, bias, out, pre, aux, colstats)
    _gemm(a)
    return (out, pre) if want_pre else out


def linear_dgrad(dy, w, epi=L.EPI_LINEAR, alpha=1.0, aux=None, out_dtype=None, drop=None):
    """dx[M,K] = epi(dy[M,N] @ w[N,K])  (B operand read transposed in place: MN-major descriptor, no copy).
    drop = (rng, p, site): the gradient of an nn.Dropout in front of this layer's input, i.e. drop(epi(..)) with the forward's mask."""
    _cuda(dy, w)
    M, N = dy.shape
    K = w.shape[1]
    out = torch.empty((M, K), device=dy.device, dtype=out_dtype or dy.dtype)
    a = L.GemmArgs()
    a.mode, a.impl, a.M, a.N, a.K = L.GEMM_PLAIN, GEMM_IMPL, M, K, N
    a.A, a.sam, a.sak = dy.data_ptr(), dy.stride(0), 1
    a.B, a.sbn, a.sbk = w.data_ptr(), 1, w.stride(0)
    a.ab_dtype = _dt(dy)
    _epi(a, epi, alpha, None, out, None, aux)
    if drop is not None and drop[1] > 0:
        if not _gemm_drop(a, drop):
            _gemm(a)
            dropout_rng(drop[0], out, drop[1], drop[2], out=out)
        return out
    _gemm(a)
    return out


def linear_wgrad(dy, x, alpha=1.0):
    """dw[N,K] (fp32) = alpha * dy[M,N]^T @ x[M,K], split-K over the M tokens with atomic fp32 accumulation."""
    _cuda(dy, x)
    M, N = dy.shape
    K = x.shape[1]
    dw = zeros_f32((N, K), dy.device)
    a = L.GemmArgs()
    a.mode, a.impl, a.M, a.N, a.K = L.GEMM_PLAIN, GEMM_IMPL, N, K, M
    a.A, a.sam, a.sak = dy.data_ptr(), 1, dy.stride(0)
    a.B, a.sbn, a.sbk = x.data_ptr(), 1, x.stride(0)
    a.ab_dtype = _dt(dy)
    _epi(a, L.EPI_ACCUM, alpha, None, dw)
    a.split_k = _split_for(N, K, M)
    _gemm(a)
    return dw


def colsum(x, alpha=1.0):
    _cuda(x)
    rows, Cn = x.shape
    out = zeros_f32((Cn,), x.device)
    L.check(L.load().avec_colsum(x.data_ptr(), _dt(x), rows, Cn, x.stride(0), alpha, out.data_ptr(), 1, _stream()), "avec_colsum")
    return out


def make_geom(N, Ti, Hi, Wi, Cin, Cout, k, s, p):
    g = L.ConvGeom()
    g.N, g.Ti, g.Hi, g.Wi, g.C, g.Co = N, Ti, Hi, Wi, Cin, Cout
    g.KT, g.KH, g.KW = k
    g.st, g.sh, g.sw = s
    g.pt, g.ph, g.pw = p
    g.To = (Ti + (k[0] - 1) - k[0]) // s[0] + 1   # "same" pre-padding: total pad = k-1  (layers.py:250-258)
    g.Ho = (Hi + (k[1] - 1) - k[1]) // s[1] + 1
    g.Wo = (Wi + (k[2] - 1) - k[2]) // s[2] + 1
    return g


def geom_sites(g, out=True):
    return g.N * g.To * g.Ho * g.Wo if out else g.N * g.Ti * g.Hi * g.Wi


def conv_fwd(x, wp, g, bias=None, epi=L.EPI_LINEAR, colstats=None, aux=None):
    """x [N,Ti,Hi,Wi,C] channels-last, wp [Co, taps*C] -> y [sites_out, Co]."""
    _cuda(x, wp)
    M = geom_sites(g, True)
    out = torch.empty((M, g.Co), device=x.device, dtype=x.dtype)
    a = L.GemmArgs()
    a.mode, a.impl, a.M, a.N, a.K = L.GEMM_CONV_FWD, GEMM_IMPL, M, g.Co, wp.shape[1]
    a.A, a.B, a.ab_dtype, a.g = x.data_ptr(), wp.data_ptr(), _dt(x), g
    _epi(a, epi, 1.0, bias, out, None, aux, colstats)
    _gemm(a)
    return out


DGRAD_CLASSES = os.environ.get("AVEC_DGRAD_CLASSES", "1") != "0"


def conv_dgrad(dy, wd, g, epi=L.EPI_LINEAR, aux=None):
    """dy [sites_out, Co], wd [C, taps*Co] -> dx [sites_in, C] (optionally + aux: merging two gradient branches)."""
    _cuda(dy, wd)
    if (g.sh > 1 or g.sw > 1) and g.sh == g.sw and g.KT == 1 and g.Ti == 1 and dy.dtype == torch.bfloat16 and GEMM_IMPL != L.IMPL_SIMT \
            and g.C % 64 == 0 and g.Co % 64 == 0 and not (g.KH == 3 and g.KW == 3 and g.sh == 2 and DGRAD_CLASSES):
        # strided 1x1 conv (ResNet shortcuts): insert zeros into dY and run the stride-1 TMA dgrad.  Stride-2 3x3 convs go
        # straight to avec_gemm, which runs one exact launch per output parity class (9 tap-GEMMs instead of 36)
        up = torch.empty((g.N * g.Hi * g.Wi, g.Co), device=dy.device, dtype=dy.dtype)
        L.check(L.load().avec_zero_upsample(dy.data_ptr(), up.data_ptr(), g.N, g.Ho, g.Wo, g.Hi, g.Wi, g.Co, g.sh, _dt(dy), _stream()),
                "avec_zero_upsample")
        g1 = make_geom(g.N, 1, g.Hi, g.Wi, g.C, g.Co, (1, g.KH, g.KW), (1, 1, 1), (0, g.ph, g.pw))
        return conv_dgrad(up, wd, g1, epi, aux)
    M = geom_sites(g, False)
    out = torch.empty((M, g.C), device=dy.device, dtype=dy.dtype)
    a = L.GemmArgs()
    a.mode, a.impl, a.M, a.N, a.K = L.GEMM_CONV_DGRAD, GEMM_IMPL, M, g.C, wd.shape[1]
    a.A, a.B, a.ab_dtype, a.g = dy.data_ptr(), wd.data_ptr(), _dt(dy), g
    _epi(a, epi, 1.0, None, out, None, aux)
    _gemm(a)
    return out


def conv_wgrad(dy, x, g):
    """dw [Co, taps*C] fp32 = sum over output sites of dy[site, co] * x[shift(site, tap), ci]."""
    _cuda(dy, x)
    taps = g.KT * g.KH * g.KW
    sites = geom_sites(g, True)
    dw = zeros_f32((g.Co, taps * g.C), dy.device)
    a = L.GemmArgs()
    a.mode, a.impl, a.M, a.N, a.K = L.GEMM_CONV_WGRAD, GEMM_IMPL, g.Co, taps * g.C, sites
    a.A, a.B, a.ab_dtype, a.g = dy.data_ptr(), x.data_ptr(), _dt(dy), g
    _epi(a, L.EPI_ACCUM, 1.0, None, dw)
    a.split_k = _split_for(g.Co, taps * g.C, sites)
    _gemm(a)
    return dw


# ------------------------------------------------------------------------------------------------------ normalisations
def layernorm_fwd(x, gamma, beta, P=1, eps=1e-6, pad_out=True):
    """x [B,T,C] -> y [B,ceil(T/P),C] (mean over P consecutive LayerNorm'ed frames), mean/rstd [B*T] fp32.  y feeds GEMMs as an
    operand: its rows get the TMA-able pitch of row_pitch() (a [B,Tp,C] view of a wider allocation when C % 8 != 0)."""
    _cuda(x)
    B, T, Cn = x.shape
    Tp = -(-T // P)
    y = empty_rows(B * Tp, Cn, x.dtype, x.device) if pad_out else torch.empty((B * Tp, Cn), device=x.device, dtype=x.dtype)
    ldy = y.stride(0)
    y3 = y.view(B, Tp, Cn)
    if gamma is None:   # identity mode: plain mean over the P frames of a patch (attention.forwardQKV without the module's norm)
        L.check(L.load().avec_layernorm_fwd(x.data_ptr(), 0, 0, y.data_ptr(), 0, 0, B, T, Cn, P, eps, _dt(x), ldy, _stream()), "avec_layernorm_fwd")
        return y3, None, None
    mean = torch.empty((B * T,), device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    L.check(L.load().avec_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                        rstd.data_ptr(), B, T, Cn, P, eps, _dt(x), ldy, _stream()), "avec_layernorm_fwd")
    return y3, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, P=1, dres=None, res_stride=1):
    """returns dx [B,T,C], dgamma, dbeta (fp32).  dx += dres[b, t/res_stride] on frames t % res_stride == 0."""
    B, T, Cn = x.shape
    dx = torch.empty_like(x)
    if gamma is None:
        L.check(L.load().avec_layernorm_bwd(dy.data_ptr(), x.data_ptr(), 0, 0, 0, _p(dres), res_stride, dx.data_ptr(), 0, 0, B, T, Cn, P,
                                            _dt(x), _stream()), "avec_layernorm_bwd")
        return dx, None, None
    dg = zeros_f32((Cn,), x.device)
    db = zeros_f32((Cn,), x.device)
    L.check(L.load().avec_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                        _p(dres), res_stride, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), B, T, Cn, P,
                                        _dt(x), _stream()), "avec_layernorm_bwd")
    return dx, dg, db


def upsample_add(x, o, P):
    B, T, Cn = x.shape
    y = torch.empty_like(x)
    L.check(L.load().avec_upsample_add(x.data_ptr(), o.data_ptr(), y.data_ptr(), B, T, o.shape[1], Cn, P, _dt(x), _stream()),
            "avec_upsample_add")
    return y


def pool_sum(dy, P):
    B, T, Cn = dy.shape
    Tp = -(-T // P)
    out = empty_rows(B * Tp, Cn, dy.dtype, dy.device)      # a GEMM operand of the backward: TMA-able row pitch
    L.check(L.load().avec_pool_sum(dy.data_ptr(), out.data_ptr(), B, T, Tp, Cn, P, _dt(dy), out.stride(0), _stream()), "avec_pool_sum")
    return out.view(B, Tp, Cn)


def softmax_fwd(x, out_dtype):
    rows, Cn = x.shape
    y = torch.empty((rows, Cn), device=x.device, dtype=out_dtype)
    L.check(L.load().avec_softmax_fwd(x.data_ptr(), _dt(x), y.data_ptr(), _dt(y), rows, Cn, _stream()), "avec_softmax_fwd")
    return y


def softmax_bwd(dy, y, dadd=None, out_dtype=None):
    rows, Cn = y.shape
    dx = torch.empty((rows, Cn), device=y.device, dtype=out_dtype or y.dtype)
    L.check(L.load().avec_softmax_bwd(dy.data_ptr(), y.data_ptr(), _dt(y), _p(dadd), dx.data_ptr(), _dt(dx), rows, Cn,
                                      _stream()), "avec_softmax_bwd")
    return dx


def bn_stats(u2d):
    rows, Cn = u2d.shape
    stats = zeros_f32((2 * Cn,), u2d.device)
    L.check(L.load().avec_bn_stats(u2d.data_ptr(), _dt(u2d), rows, Cn, stats.data_ptr(), _stream()), "avec_bn_stats")
    return stats


def gemm_stats_buffer(Cn, device):
    """zeroed accumulator for the GEMM-epilogue BatchNorm statistics: [STATS_REPLICAS][2*C]"""
    return zeros_f32((L.STATS_REPLICAS * 2 * Cn,), device)


# SyncBatchNorm switch (the reference's DDP default, nnet/model.py:59-61): when set, the per-channel sums of every training-mode
# BatchNorm are all-reduced over the data-parallel group - forward [sum x, sum x^2] (one packed all-reduce of the accumulator
# the producing kernel filled), backward [sum dy, sum dy*xhat] - so all ranks normalise with the statistics of the GLOBAL batch.
# Default None = local statistics (the reference's sync_batch_norm=False branch, model.py:62-63; what north_star prescribes).
SYNC_BN = None


def set_sync_batchnorm(enabled, group=None):
    global SYNC_BN
    import torch.distributed as dist
    SYNC_BN = (group if group is not None else dist.group.WORLD) if enabled else None


def _sync_world():
    import torch.distributed as dist
    return dist.get_world_size(SYNC_BN)


def bn_finalize(stats, gamma, beta, count, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    Cn = gamma.numel() if gamma is not None else stats.numel() // 2
    replicas = stats.numel() // (2 * Cn)
    if SYNC_BN is not None:
        import torch.distributed as dist
        dist.all_reduce(stats, group=SYNC_BN)
        count = count * _sync_world()
    buf = torch.empty((4, Cn), device=stats.device, dtype=torch.float32)  # scale, shift, mean, rstd
    L.check(L.load().avec_bn_finalize(stats.data_ptr(), _p(gamma), _p(beta), buf[0].data_ptr(), buf[1].data_ptr(),
                                      buf[2].data_ptr(), buf[3].data_ptr(), _p(running_mean), _p(running_var), count, Cn,
                                      eps, momentum, replicas, _stream()), "avec_bn_finalize")
    return buf


def bn_eval_affine(gamma, beta, running_mean, running_var, eps=1e-5):
    Cn = running_mean.numel()
    buf = torch.empty((2, Cn), device=running_mean.device, dtype=torch.float32)
    L.check(L.load().avec_bn_eval_affine(_p(gamma), _p(beta), running_mean.data_ptr(), running_var.data_ptr(),
                                         buf[0].data_ptr(), buf[1].data_ptr(), Cn, eps, _stream()), "avec_bn_eval_affine")
    return buf


def bn_apply(u2d, scale, shift, act, res=None, pad_out=False):
    """pad_out: y feeds a GEMM as an operand (ConvModule: the pointwise conv after BatchNorm + Swish): TMA-able row pitch"""
    rows, Cn = u2d.shape
    y = empty_rows(rows, Cn, u2d.dtype, u2d.device) if pad_out else torch.empty_like(u2d)
    L.check(L.load().avec_bn_apply(u2d.data_ptr(), scale.data_ptr(), shift.data_ptr(), _p(res), y.data_ptr(), rows, Cn, act,
                                   _dt(u2d), y.stride(0), _stream()), "avec_bn_apply")
    return y


def bn_bwd(dy2d, u2d, bnbuf, gamma, act, res=None, want_dres=False):
    """BatchNorm (+activation, + residual add before it) backward with batch statistics.
    returns du, dres (or None), dgamma, dbeta."""
    rows, Cn = u2d.shape
    sums = zeros_f32((2 * Cn,), u2d.device)
    lib = L.load()
    L.check(lib.avec_bn_bwd_reduce(dy2d.data_ptr(), u2d.data_ptr(), bnbuf[0].data_ptr(), bnbuf[1].data_ptr(), _p(res),
                                   bnbuf[2].data_ptr(), bnbuf[3].data_ptr(), sums.data_ptr(), rows, Cn, act, _dt(u2d),
                                   _stream()), "avec_bn_bwd_reduce")
    du = torch.empty_like(u2d)
    dres = torch.empty_like(u2d) if want_dres else None
    gsums = sums
    if SYNC_BN is not None:
        # the kernel divides by the local row count: hand it the MEAN over ranks of the sums (= global sums / global count);
        # dgamma / dbeta stay the local sums, the gradient all-reduce averages them like every other parameter gradient
        import torch.distributed as dist
        gsums = sums.clone()
        dist.all_reduce(gsums, group=SYNC_BN)
        gsums.div_(_sync_world())
    L.check(lib.avec_bn_bwd_apply(dy2d.data_ptr(), u2d.data_ptr(), bnbuf[0].data_ptr(), bnbuf[1].data_ptr(), _p(res),
                                  bnbuf[2].data_ptr(), bnbuf[3].data_ptr(), _p(gamma), gsums.data_ptr(), du.data_ptr(),
                                  _p(dres), rows, Cn, act, _dt(u2d), _stream()), "avec_bn_bwd_apply")
    return du, dres, sums[Cn:], sums[:Cn]


def bn_bwd_pool(dyp, idx, u2d, bnbuf, gamma, N, Hi, Wi):
    """BatchNorm backward fed by the max-pool backward (gathered on the fly): returns du, dgamma, dbeta"""
    _cuda(dyp, u2d)
    Cn = u2d.shape[1]
    Ho, Wo = (Hi - 1) // 2 + 1, (Wi - 1) // 2 + 1
    sums = zeros_f32((2 * Cn,), u2d.device)
    du = torch.empty_like(u2d)
    L.check(L.load().avec_bn_bwd_pool(dyp.data_ptr(), idx.data_ptr(), u2d.data_ptr(), bnbuf[2].data_ptr(), bnbuf[3].data_ptr(), _p(gamma),
                                      sums.data_ptr(), du.data_ptr(), N, Hi, Wi, Cn, Ho, Wo, _dt(u2d), _stream()), "avec_bn_bwd_pool")
    return du, sums[Cn:], sums[:Cn]


# ----------------------------------------------------------------------------------------------------------- attention
def relpos_attn_fwd(qkv, e, klen, qlen, B, T, H, d, G=1, Tf=None, u=None, v=None):
    """qkv [B*Tf, 3*D1] (frames), e [2T-1, G*D1] -> o [B*Tf, D1], probs [B,H,T,T]; T = ceil(Tf/G) tokens of G frames"""
    Tf = T if Tf is None else Tf
    D1 = H * d // G
    o = torch.empty((B * Tf, D1), device=qkv.device, dtype=qkv.dtype)
    probs = torch.empty((B, H, T, T), device=qkv.device, dtype=torch.float32)
    L.check(L.load().avec_relpos_attn_fwd(qkv.data_ptr(), e.data_ptr(), _p(klen), qlen, o.data_ptr(), probs.data_ptr(), B, T,
                                          H, d, G, Tf, _p(u), _p(v), _dt(qkv), _stream()), "avec_relpos_attn_fwd")
    return o, probs


def relpos_attn_bwd(do, qkv, e, probs, B, T, H, d, G=1, Tf=None, u=None, v=None):
    Tf = T if Tf is None else Tf
    D1 = H * d // G
    dqkv = torch.empty_like(qkv)
    de = zeros_f32((2 * T - 1, G * D1), qkv.device)
    du = zeros_f32((D1,), qkv.device) if u is not None else None
    dv = zeros_f32((D1,), qkv.device) if v is not None else None
    ws = torch.empty_like(probs)
    L.check(L.load().avec_relpos_attn_bwd(do.data_ptr(), qkv.data_ptr(), e.data_ptr(), probs.data_ptr(), ws.data_ptr(),
                                          dqkv.data_ptr(), de.data_ptr(), B, T, H, d, G, Tf, _p(u), _p(v), _p(du), _p(dv),
                                          _dt(qkv), _stream()), "avec_relpos_attn_bwd")
    return dqkv, de, du, dv


# tcgen05 / TMEM / TMA attention (csrc/attention_tc.cu): bf16, regular + patch attention, any sequence length, flash-style
ATTN_TC = os.environ.get("AVEC_ATTN_TC", "1") != "0"


def set_attention_tc(enabled):
    """False: the round-1 kernels (mma.sync up to 128 keys, SIMT beyond) - kept for grouped attention, fp32 and A/B tests"""
    global ATTN_TC
    ATTN_TC = bool(enabled)


def attn_head_pad(d):
    """column block of one head in the padded-heads layout (one or two 64-wide TMA boxes), None if unsupported"""
    return 64 if d <= 64 else (128 if d <= 128 else (192 if d <= 192 else (256 if d <= 256 else None)))


def attn_group_pack(src, u, v, B, Tf, Tn, G, H, D1, dp, parts):
    """frame-rate [B*Tf, (3*)D1] -> padded-heads tokens [B*Tn, parts*H*dp] bf16 (parts 4: q+u | k | v | q+v; 1: plain regroup)"""
    _cuda(src)
    dst = torch.empty((B * Tn, parts * H * dp), device=src.device, dtype=torch.bfloat16)
    L.check(L.load().avec_attn_group_pack(src.data_ptr(), _dt(src), src.stride(0), _p(u), _p(v), dst.data_ptr(), dst.stride(0), B, Tf, Tn, G, H,
                                          D1, dp, parts, _stream()), "avec_attn_group_pack")
    return dst


def attn_group_unpack(src, B, Tf, Tn, G, H, D1, dp):
    """padded-heads tokens [B*Tn, H*dp] -> frames [B*Tf, D1] (same dtype: bf16 o, fp32 de)"""
    _cuda(src)
    dst = torch.empty((B * Tf, D1), device=src.device, dtype=src.dtype)
    L.check(L.load().avec_attn_group_unpack(src.data_ptr(), _dt(src), src.stride(0), dst.data_ptr(), _dt(dst), dst.stride(0), B, Tf, Tn, G, H, D1,
                                            dp, _stream()), "avec_attn_group_unpack")
    return dst


def attn_group_unpack_dqkv(dqkv_tok, B, Tf, Tn, G, H, D1, dp, want_bias=True):
    """[B*Tn, 4*H*dp] -> dqkv frames [B*Tf, 3*D1] bf16, du, dv [D1] fp32"""
    _cuda(dqkv_tok)
    dst = torch.empty((B * Tf, 3 * D1), device=dqkv_tok.device, dtype=dqkv_tok.dtype)
    du = zeros_f32((D1,), dqkv_tok.device) if want_bias else None
    dv = zeros_f32((D1,), dqkv_tok.device) if want_bias else None
    L.check(L.load().avec_attn_group_unpack_dqkv(dqkv_tok.data_ptr(), dqkv_tok.stride(0), dst.data_ptr(), dst.stride(0), _p(du), _p(dv), B, Tf, Tn,
                                                 G, H, D1, dp, _stream()), "avec_attn_group_unpack_dqkv")
    return dst, du, dv


def unpad_heads(w, H, d, dp, cols=False):
    """fp32 gradient in the padded-heads layout -> dense: rows [n*H*dp, K] -> [n*H*d, K] (or 1-d), cols [N, H*dp] -> [N, H*d]"""
    _cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous()
    if cols:
        N = w.shape[0]
        out = torch.empty((N, H * d), device=w.device, dtype=torch.float32)
        L.check(L.load().avec_unpad_heads(w.data_ptr(), out.data_ptr(), N * H, d, dp, 1, _stream()), "avec_unpad_heads")
        return out
    K = w.numel() // w.shape[0]
    groups = w.shape[0] // dp
    out = torch.empty((groups * d,) + tuple(w.shape[1:]), device=w.device, dtype=torch.float32)
    L.check(L.load().avec_unpad_heads(w.data_ptr(), out.data_ptr(), groups, d, dp, K, _stream()), "avec_unpad_heads")
    return out


def relpos_attn_tc_fwd(qkv, e, klen, qlen, B, T, H, d, dp, qp_part=0):
    """qkv [B*T, 3*H*dp] (qp_part = 3: [B*T, 4*H*dp] = q+u | k | v | q+v), e [2T-1, H*dp] (padded heads) -> o [B*T, H*dp] bf16, lse [B,H,T]"""
    _cuda(qkv, e)
    o = torch.empty((B * T, H * dp), device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32)
    L.check(L.load().avec_relpos_attn_tc_fwd(qkv.data_ptr(), qkv.stride(0), e.data_ptr(), e.stride(0), _p(klen), int(qlen), o.data_ptr(),
                                             o.stride(0), lse.data_ptr(), B, T, H, d, dp, int(qp_part), _stream()), "avec_relpos_attn_tc_fwd")
    return o, lse


def relpos_attn_tc_bwd(do, qkv, e, o, lse, klen, qlen, B, T, H, d, dp, qp_part=0):
    """-> dqkv (qkv's layout) bf16, de [2T-1, H*dp] fp32"""
    _cuda(do, qkv, e, o)
    nparts = 4 if qp_part else 3
    dqkv = torch.empty((B * T, nparts * H * dp), device=qkv.device, dtype=qkv.dtype)
    ws = None if T <= 128 else zeros_f32((B * T, nparts * H * dp), qkv.device)
    de = zeros_f32((2 * T - 1, H * dp), qkv.device)
    L.check(L.load().avec_relpos_attn_tc_bwd(do.data_ptr(), do.stride(0), qkv.data_ptr(), qkv.stride(0), e.data_ptr(), e.stride(0),
                                             o.data_ptr(), o.stride(0), lse.data_ptr(), _p(klen), int(qlen), dqkv.data_ptr(), dqkv.stride(0),
                                             _p(ws), de.data_ptr(), de.stride(0), B, T, H, d, dp, int(qp_part), _stream()), "avec_relpos_attn_tc_bwd")
    if ws is not None:
        dqkv = convert(ws, qkv.dtype)
    return dqkv, de


# --------------------------------------------------------------------------------------------------------- conv module
def glu_dwconv_fwd(pre, w, bias, stride, ksize=15, want_stats=True):
    B, T, C2 = pre.shape
    Cn = C2 // 2
    pad = (ksize - 1) // 2
    To = (T + 2 * pad - ksize) // stride + 1
    u = torch.empty((B, To, Cn), device=pre.device, dtype=pre.dtype)
    stats = zeros_f32((2 * Cn,), pre.device) if want_stats else None
    L.check(L.load().avec_glu_dwconv_fwd(pre.data_ptr(), w.data_ptr(), _p(bias), u.data_ptr(), _p(stats), B, T, To, Cn, ksize,
                                         stride, pad, _dt(pre), _stream()), "avec_glu_dwconv_fwd")
    return u, stats


def glu_dwconv_bwd(du, pre, w, stride, ksize=15):
    B, T, C2 = pre.shape
    Cn = C2 // 2
    pad = (ksize - 1) // 2
    To = du.shape[1]
    dpre = torch.empty_like(pre)
    dw = zeros_f32((Cn, ksize), pre.device)
    db = zeros_f32((Cn,), pre.device)
    L.check(L.load().avec_glu_dwconv_bwd(du.data_ptr(), pre.data_ptr(), w.data_ptr(), dpre.data_ptr(), dw.data_ptr(),
                                         db.data_ptr(), B, T, To, Cn, ksize, stride, pad, _dt(pre), _stream()),
            "avec_glu_dwconv_bwd")
    return dpre, dw, db


# ---------------------------------------------------------------------------------------------------------- front-ends
def stft_mel_log(wave, fb, layout=0):
    _cuda(wave, fb)
    B, Ln = wave.shape
    F = Ln // 160 + 1
    out = torch.empty((B, F, 80) if layout == 0 else (B, 80, F), device=wave.device, dtype=torch.float32)
    L.check(L.load().avec_stft_mel_log(wave.data_ptr(), fb.data_ptr(), out.data_ptr(), B, Ln, F, layout, _stream()),
            "avec_stft_mel_log")
    return out


def im2col_c1(x, g, Kpad):
    """x [N,Ti,Hi,Wi,1] -> col [sites_out, Kpad] (filter taps, zero padded)"""
    col = torch.empty((geom_sites(g), Kpad), device=x.device, dtype=x.dtype)
    L.check(L.load().avec_im2col_c1(x.data_ptr(), col.data_ptr(), g, Kpad, _dt(x), _stream()), "avec_im2col_c1")
    return col


def stem2d_fwd(x, w9, bias, colstats=None):
    """audio stem conv: x [N,H,W] (C = 1), w9 [Co, 9] -> u [N*Ho*Wo, Co] (+ BatchNorm column sums)"""
    _cuda(x, w9)
    N, H, W = x.shape[0], x.shape[1], x.shape[2]
    Co = w9.shape[0]
    out = torch.empty((N * ((H - 1) // 2 + 1) * ((W - 1) // 2 + 1), Co), device=x.device, dtype=x.dtype)
    L.check(L.load().avec_stem2d_fwd(x.data_ptr(), w9.data_ptr(), _p(bias), out.data_ptr(), _p(colstats), N, H, W, Co, _dt(x), _stream()), "avec_stem2d_fwd")
    return out


def stem2d_wgrad(x, dy):
    _cuda(x, dy)
    N, H, W = x.shape[0], x.shape[1], x.shape[2]
    Co = dy.shape[1]
    dw = zeros_f32((Co, 9), x.device)
    L.check(L.load().avec_stem2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, Co, _dt(x), _stream()), "avec_stem2d_wgrad")
    return dw


STEM_KPAD = 320   # forward weight row of the direct stem kernel: k = (kt*7+kh)*8 + kw, zero padded (ST_KPAD in gemm_tc.cu)


def stem3d_supported(x, Co, kt, kh, kw):
    """the direct tcgen05 stem kernel covers the reference's visual stem: bf16, 1 -> 64 channels, k (5,7,7), s (1,2,2), 88-pixel rows
    (a 128-site tile must span at most 4 output rows: its staged input window holds 13 rows)"""
    B, T, H, W = x.shape[0], x.shape[1], x.shape[2], x.shape[3]
    return (x.dtype == torch.bfloat16 and Co == 64 and (kt, kh, kw) == (5, 7, 7) and W == 88 and H % 2 == 0 and H >= 8
            and GEMM_IMPL != L.IMPL_SIMT)


def stem3d_weight_layout(cw):
    """kernel layout of the visual stem filter: (64, 1, 5, 7, 7) -> bf16 [64, 320] with k = (kt*7+kh)*8 + kw (kw = 7 and k >= 280 zero)"""
    from . import weights as W
    return W.PLAN.layout((cw,), "stem3d_direct", W.b_custom(lambda w: w[:, 0].reshape(w.shape[0], 35, 7), cw.shape[0], STEM_KPAD, (STEM_KPAD, 8, 1)),
                         torch.bfloat16)


def stem3d_pack_weight(w):
    """(64, 1, 5, 7, 7) -> [64, 320] with k = (kt*7+kh)*8 + kw"""
    Co = w.shape[0]
    wp = torch.zeros((Co, 35, 8), device=w.device, dtype=w.dtype)
    wp[:, :, :7] = w.reshape(Co, 35, 7)
    return torch.nn.functional.pad(wp.reshape(Co, 280), (0, STEM_KPAD - 280))


def stem3d_fwd(x, wp, bias, colstats=None):
    """x [B,T,H,W,1] bf16, wp [64,320] -> u [B*T*(H/2)*(W/2), 64] (+ BatchNorm column sums)"""
    _cuda(x, wp)
    B, T, H, W = x.shape[0], x.shape[1], x.shape[2], x.shape[3]
    out = torch.empty((B * T * (H // 2) * (W // 2), 64), device=x.device, dtype=x.dtype)
    L.check(L.load().avec_stem3d_fwd(x.data_ptr(), wp.data_ptr(), _p(bias), out.data_ptr(), _p(colstats), B, T, H, W, _stream()), "avec_stem3d_fwd")
    return out


def stem3d_wgrad(x, dy):
    """dw [64, 245] fp32 of the same convolution (x [B,T,H,W,1] bf16, dy [sites, 64] bf16)"""
    _cuda(x, dy)
    B, T, H, W = x.shape[0], x.shape[1], x.shape[2], x.shape[3]
    dw = zeros_f32((64, 245), x.device)
    L.check(L.load().avec_stem3d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, T, H, W, _stream()), "avec_stem3d_wgrad")
    return dw


def bn_relu_maxpool_fwd(u, scale, shift, N, Hi, Wi, Cn):
    Ho, Wo = (Hi - 1) // 2 + 1, (Wi - 1) // 2 + 1
    y = torch.empty((N, Ho, Wo, Cn), device=u.device, dtype=u.dtype)
    idx = torch.empty((N, Ho, Wo, Cn), device=u.device, dtype=torch.uint8)
    L.check(L.load().avec_bn_relu_maxpool_fwd(u.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), idx.data_ptr(), N,
                                              Hi, Wi, Cn, Ho, Wo, _dt(u), _stream()), "avec_bn_relu_maxpool_fwd")
    return y, idx


def bn_relu_maxpool_bwd(dy, idx, N, Hi, Wi, Cn):
    Ho, Wo = idx.shape[1], idx.shape[2]
    dz = torch.empty((N * Hi * Wi, Cn), device=dy.device, dtype=dy.dtype)
    L.check(L.load().avec_bn_relu_maxpool_bwd(dy.data_ptr(), idx.data_ptr(), dz.data_ptr(), N, Hi, Wi, Cn, Ho, Wo, _dt(dy),
                                              _stream()), "avec_bn_relu_maxpool_bwd")
    return dz


def avgpool_fwd(x, N, HW, Cn):
    y = torch.empty((N, Cn), device=x.device, dtype=x.dtype)
    L.check(L.load().avec_avgpool_fwd(x.data_ptr(), y.data_ptr(), N, HW, Cn, _dt(x), _stream()), "avec_avgpool_fwd")
    return y


def avgpool_bwd(dy, N, HW, Cn):
    dx = torch.empty((N * HW, Cn), device=dy.device, dtype=dy.dtype)
    L.check(L.load().avec_avgpool_bwd(dy.data_ptr(), dx.data_ptr(), N, HW, Cn, _dt(dy), _stream()), "avec_avgpool_bwd")
    return dx


def ctc_loss(logits, labels, in_len, lab_len, blank=0, zero_infinity=False):
    """logits [B,T,V] fp32 -> (nll [B], grad [B,T,V]); lengths are int64 device tensors (no host sync)."""
    _cuda(logits, labels, lab_len)
    B, T, V = logits.shape
    Lmax = labels.shape[1]
    nll = torch.empty((B,), device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits)
    ws = torch.empty((B * T * (2 * Lmax + 1),), device=logits.device, dtype=torch.float32)
    L.check(L.load().avec_ctc_loss(logits.data_ptr(), labels.data_ptr(), _p(in_len), lab_len.data_ptr(), nll.data_ptr(), grad.data_ptr(),
                                   ws.data_ptr(), B, T, V, Lmax, blank, 1 if zero_infinity else 0, _stream()), "avec_ctc_loss")
    return nll, grad


def row_pitch(C, dtype):
    """leading dimension (elements) of a [rows, C] activation / weight matrix: rows start on 16-byte boundaries so that
    every operand qualifies for a TMA descriptor (D = 180 bf16 rows are 360 B: pitch 184 -> 192 keeps them on 128-byte lines)"""
    if dtype == torch.bfloat16 and C % 8 != 0:
        return (C + 63) // 64 * 64
    return C


def empty_rows(rows, C, dtype, device):
    """[rows, C] view of a [rows, row_pitch(C)] allocation (contiguous when the pitch equals C)"""
    ld = row_pitch(C, dtype)
    if ld == C:
        return torch.empty((rows, C), device=device, dtype=dtype)
    return torch.empty((rows, ld), device=device, dtype=dtype)[:, :C]


def convert(src, dtype, pad=False):
    """dtype cast through the library's own kernel (2-d, possibly strided rows).  pad: the destination rows get the TMA-able
    pitch of row_pitch() (a [rows, C] view of a wider allocation)."""
    if src.dtype == dtype and src.is_contiguous() and not (pad and src.dim() == 2 and row_pitch(src.shape[1], dtype) != src.shape[1]):
        return src
    s2 = src.reshape(-1, src.shape[-1]) if src.dim() != 2 else src
    if s2.stride(-1) != 1:
        s2 = s2.contiguous()
    if pad and src.dim() == 2:
        dst = empty_rows(s2.shape[0], s2.shape[1], dtype, src.device)
    else:
        dst = torch.empty(s2.shape, device=src.device, dtype=dtype)
    L.check(L.load().avec_convert(s2.data_ptr(), _dt(s2), s2.stride(0), dst.data_ptr(), _dt(dst), dst.stride(0), s2.shape[0],
                                  s2.shape[1], _stream()), "avec_convert")
    return dst if (pad and src.dim() == 2) else dst.reshape(src.shape)


# ------------------------------------------------------------------------------------- training-step kernels (train.cu)
class _Rng:
    """{seed, step} as two uint64 in device memory per GPU, plus the per-forward dropout site counter.  The step is advanced
    by a kernel (capturable: every CUDA-graph replay draws fresh masks); sites are handed out in call order and are
    therefore identical in the forward and its backward.  Every forward pass works on a SNAPSHOT of {seed, step} taken when
    the pass starts (functional.new_step), which its autograd Functions keep: 'forward A, forward B, backward A' regenerates
    A's masks.  Under torch.distributed the rank is folded into the seed so that replicas draw different masks."""

    GOLDEN = 0x9E3779B97F4A7C15

    def __init__(self):
        self.state, self.snap, self.seed, self.site = {}, {}, 0x5EEDA7EC, 0

    def _rank_seed(self):
        rank = 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank = dist.get_rank()
        except Exception:  # noqa: BLE001
            rank = 0
        return (self.seed + rank * self.GOLDEN) & 0x7FFFFFFFFFFFFFFF

    def manual_seed(self, seed):
        self.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        for st in self.state.values():
            st.copy_(torch.tensor([self._rank_seed(), 0], dtype=torch.int64))
        self.snap.clear()

    @staticmethod
    def _dev(device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return device

    def get(self, device):
        device = self._dev(device)
        st = self.state.get(device)
        if st is None:
            st = torch.tensor([self._rank_seed(), 0], dtype=torch.int64).to(device)
            self.state[device] = st
        return st

    def advance(self, device):
        st = self.get(device)
        L.check(L.load().avec_counter_advance(st.data_ptr() + 8, _stream()), "avec_counter_advance")

    def snapshot(self, device):
        """copy of the live {seed, step} for the forward pass that starts now (one 16-byte device copy, capturable)"""
        device = self._dev(device)
        self.snap[device] = self.get(device).clone()
        return self.snap[device]

    def cur(self, device):
        device = self._dev(device)
        s = self.snap.get(device)
        return s if s is not None else self.get(device)

    def next_site(self):
        self.site += 1
        return self.site


RNG = _Rng()


def dropout(x, p, site, res=None, alpha=1.0, out=None, up=None, pad_out=False):
    """dropout_rng on the live generator state (tests / direct calls)"""
    return dropout_rng(RNG.get(x.device), x, p, site, res, alpha, out, up, pad_out)


def dropout_rng(rng, x, p, site, res=None, alpha=1.0, out=None, up=None, pad_out=False):
    """out = (res or 0) + alpha * keep * x / (1 - p) with the Philox mask of (rng = {seed, step} device tensor, site); the
    backward calls it again on the gradient with the same rng / site.  up = (T, Tp, P): x is [B*Tp, C] patch rows repeated over
    the T frames of out / res.  rng None: the current forward pass's snapshot."""
    _cuda(x, res)
    rng = RNG.cur(x.device) if rng is None else rng
    C = x.shape[-1]
    if up is None:
        rows, T, Tp, P = x.numel() // C, 0, 0, 1
    else:
        T, Tp, P = up
        rows = (x.numel() // C // Tp) * T
    if out is None:     # pad_out (backward: the masked gradient is a GEMM operand): [rows, C] view with the TMA-able row pitch
        out = empty_rows(rows, C, x.dtype, x.device) if pad_out else (torch.empty_like(x) if up is None else torch.empty((rows, C), device=x.device, dtype=x.dtype))
    ldy = out.stride(-2) if out.dim() >= 2 else C
    assert x.is_contiguous() and (out.is_contiguous() or (out.dim() == 2 and out.stride(1) == 1)) and (res is None or (res.is_contiguous() and res.dtype == x.dtype))
    L.check(L.load().avec_dropout(x.data_ptr(), _p(res), out.data_ptr(), rows, C, _dt(x), float(p), float(alpha),
                                  rng.data_ptr(), int(site), T, Tp, P, ldy, _stream()), "avec_dropout")
    return out


def spec_augment_(mel, lengths, site, mF=2, Fmax=27, mT=5, pS=0.05, want_intervals=False, rng=None):
    """in-place SpecAugment of mel [B, F, M] fp32 (frame-major); lengths [B] int64 device tensor of valid frames or None"""
    _cuda(mel, lengths)
    assert mel.dtype == torch.float32 and mel.is_contiguous()
    B, F, M = mel.shape
    iv = torch.empty((B, mF + mT, 2), device=mel.device, dtype=torch.int32) if want_intervals else None
    L.check(L.load().avec_spec_augment(mel.data_ptr(), _p(lengths), B, F, M, mF, Fmax, mT, float(pS),
                                       (rng if rng is not None else RNG.cur(mel.device)).data_ptr(), int(site), _p(iv), _stream()),
            "avec_spec_augment")
    return iv


VIDEO_MAX_MASKS = 32


def video_augment(video, lengths, site, crop=(88, 88), flip_p=0.5, mask_T=10, fps=25.0, num_mask_second=1.0, want_draws=False, rng=None):
    """video [B,T,Hi,Wi] fp32 -> [B,T,Ho,Wo] fp32: RandomCrop + RandomHorizontalFlip + TimeMaskSecond per sample (see
    include/avec_b200.h); lengths [B] int64 device tensor of valid frames or None"""
    _cuda(video, lengths)
    assert video.dtype == torch.float32 and video.is_contiguous() and video.dim() == 4
    B, T, Hi, Wi = video.shape
    out = torch.empty((B, T, crop[0], crop[1]), device=video.device, dtype=torch.float32)
    fsum = torch.empty((B * T,), device=video.device, dtype=torch.float32)
    draws = torch.zeros((B, 3 + 2 * VIDEO_MAX_MASKS), device=video.device, dtype=torch.int32) if want_draws else None
    L.check(L.load().avec_video_augment(video.data_ptr(), _p(lengths), out.data_ptr(), fsum.data_ptr(), B, T, Hi, Wi, crop[0], crop[1],
                                        float(flip_p), int(mask_T), float(fps), float(num_mask_second),
                                        (rng if rng is not None else RNG.cur(video.device)).data_ptr(), int(site), _p(draws), _stream()),
            "avec_video_augment")
    return (out, draws) if want_draws else out


def ctc_greedy_decode(logits, in_len=None, blank=0, want_align=False):
    """logits [B,T,V] fp32 -> (tokens [B,T] int32 padded with -1, ntok [B] int32[, align [B,T] int32])"""
    _cuda(logits, in_len)
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    B, T, V = logits.shape
    tokens = torch.empty((B, T), device=logits.device, dtype=torch.int32)
    ntok = torch.empty((B,), device=logits.device, dtype=torch.int32)
    align = torch.empty((B, T), device=logits.device, dtype=torch.int32) if want_align else None
    L.check(L.load().avec_ctc_greedy_decode(logits.data_ptr(), _p(in_len), _p(align), tokens.data_ptr(), ntok.data_ptr(), B, T, V,
                                            blank, _stream()), "avec_ctc_greedy_decode")
    return (tokens, ntok, align) if want_align else (tokens, ntok)


def sumsq(g, out):
    """out[0] += sum g^2 over a flat fp32 buffer (out: zero-initialised device float)"""
    _cuda(g, out)
    L.check(L.load().avec_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), _stream()), "avec_sumsq")
    return out


def adam_step(p, g, m, v, step, lr_mode, lr_a, lr_b=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, sumsq_buf=None,
              max_norm=0.0, ema=None, ema_tau=0.0, info=None):
    """fused Adam over flat fp32 buffers (see include/avec_b200.h); step: device int64 holding the 1-based step count"""
    _cuda(p, g, m, v, step)
    assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32 and p.numel() == g.numel() == m.numel() == v.numel()
    L.check(L.load().avec_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(ema), p.numel(), betas[0], betas[1],
                                    eps, weight_decay, lr_mode, lr_a, lr_b, max_norm, ema_tau, step.data_ptr(), _p(sumsq_buf),
                                    _p(info), _stream()), "avec_adam_step")
