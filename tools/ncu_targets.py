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
