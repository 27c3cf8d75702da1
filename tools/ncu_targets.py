"""A handful of representative avec_gemm launches for `ncu --set full -k regex:gemm_tc` (one warm-up + one profiled each)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops, _lib as L

dev, bf = "cuda", torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def rep(fn, n=2):
    for _ in range(n):
        fn()
    torch.cuda.synchronize()


if which in ("all", "small"):     # conformer stage-2 out-projection with residual epilogue
    x, w, b = torch.randn(6464, 256, device=dev, dtype=bf), torch.randn(256, 256, device=dev, dtype=bf), torch.randn(256, device=dev)
    aux = torch.randn(6464, 256, device=dev, dtype=bf)
    rep(lambda: ops.linear_fwd(x, w, b, L.EPI_RESIDUAL, aux=aux))
if which in ("all", "conv1"):     # ResNet stage-1 conv forward with BatchNorm statistics
    N, H, W, C = 6464, 22, 22, 64
    x = torch.randn(N, H, W, C, device=dev, dtype=bf)
    g = ops.make_geom(N, 1, H, W, C, C, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    wp = torch.randn(C, 9 * C, device=dev, dtype=bf)
    st = torch.zeros(32 * 2 * C, device=dev)
    rep(lambda: ops.conv_fwd(x, wp, g, colstats=st))
    dy = torch.randn(N * H * W, C, device=dev, dtype=bf)
    rep(lambda: ops.conv_wgrad(dy, x, g))
if which in ("all", "conv3"):     # ResNet stage-3 conv forward
    N, H, W, C = 6464, 6, 6, 256
    x = torch.randn(N, H, W, C, device=dev, dtype=bf)
    g = ops.make_geom(N, 1, H, W, C, C, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    wp = torch.randn(C, 9 * C, device=dev, dtype=bf)
    rep(lambda: ops.conv_fwd(x, wp, g))
if which in ("all", "big"):
    x, w = torch.randn(8192, 8192, device=dev, dtype=bf), torch.randn(8192, 8192, device=dev, dtype=bf)
    rep(lambda: ops.linear_fwd(x, w))

if which in ("all", "attn"):      # tcgen05 flash attention at the stage-2 / video shape (B = 64, T = 101, H = 4, d = 64) and a long sequence
    for B, T, H, d in ((64, 101, 4, 64), (8, 400, 4, 64)):
        dp = ops.attn_head_pad(d)
        qkv = (0.5 * torch.randn(B * T, 3 * H * dp, device=dev)).to(bf)
        e = (0.5 * torch.randn(2 * T - 1, H * dp, device=dev)).to(bf)
        go = torch.randn(B * T, H * dp, device=dev).to(bf)
        klen = torch.full((B,), T, dtype=torch.int32, device=dev)
        o, lse = ops.relpos_attn_tc_fwd(qkv, e, klen, T, B, T, H, d, dp)
        rep(lambda: ops.relpos_attn_tc_fwd(qkv, e, klen, T, B, T, H, d, dp), 1)
        rep(lambda: ops.relpos_attn_tc_bwd(go, qkv, e, o, lse, klen, T, B, T, H, d, dp), 2)
if which in ("all", "norm"):      # the HBM-bound passes of the step: BatchNorm backward (ResNet stage 1), LayerNorm backward, dropout
    rows, C = 6464 * 22 * 22, 64
    u, dy = torch.randn(rows, C, device=dev).to(bf), torch.randn(rows, C, device=dev).to(bf)
    gamma = torch.ones(C, device=dev)
    stats = ops.bn_stats(u)
    bnbuf = ops.bn_finalize(stats, gamma, torch.zeros(C, device=dev), rows)
    rep(lambda: ops.bn_bwd(dy, u, bnbuf, gamma, L.ACT_RELU), 2)
    rep(lambda: ops.bn_apply(u, bnbuf[0], bnbuf[1], L.ACT_RELU), 2)
    B, T, D = 64, 101, 256
    x, g2 = torch.randn(B, T, D, device=dev).to(bf), torch.randn(B, T, D, device=dev).to(bf)
    w = torch.ones(D, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, w, torch.zeros(D, device=dev))
    rep(lambda: ops.layernorm_bwd(g2, x, w, mean, rstd, dres=g2), 2)
    h = torch.randn(B * T, 4 * D, device=dev).to(bf)
    rep(lambda: ops.dropout(h, 0.1, 3, out=h), 2)
