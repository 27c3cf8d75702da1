"""GPU parity of avec_gemm: SIMT fp32 against torch fp32, tcgen05 bf16 against fp32 math on the same bf16-rounded
operands (so the only difference is accumulation order: tolerance 2e-3 relative to the output scale)."""
import pytest
import torch
import torch.nn.functional as F

from avec_b200 import ops, _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, want, tol):
    scale = want.float().abs().max().item() + 1e-6
    err = (got.float() - want.float()).abs().max().item()
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


def _rand(*shape, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV).to(dtype)


SHAPES = [(300, 180, 720), (257, 720, 180), (128, 256, 1024), (1000, 360, 1440), (70, 540, 180), (513, 256, 7200), (64, 256, 512)]


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 1e-5), ("simt", torch.bfloat16, 1e-2), ("tcgen05", torch.bfloat16, 1e-2),
                                            ("tcgen05-gather", torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_fwd_dgrad_wgrad(impl, dtype, tol, M, K, N):
    ops.set_tma(impl != "tcgen05-gather")
    ops.set_gemm_impl(impl.split("-")[0])
    try:
        x, w, b = _rand(M, K, dtype=dtype, seed=1), _rand(N, K, dtype=dtype, seed=2) / K ** 0.5, _rand(N, seed=3)
        dy = _rand(M, N, dtype=dtype, seed=4)
        xf, wf, dyf = x.float(), w.float(), dy.float()
        y = ops.linear_fwd(x, w, b)
        _close(y, xf @ wf.t() + b, tol)
        h, pre = ops.linear_fwd(x, w, b, L.EPI_SWISH, want_pre=True)
        _close(pre, xf @ wf.t() + b, tol)
        _close(h, F.silu(xf @ wf.t() + b), tol)
        aux = _rand(M, N, dtype=dtype, seed=5)
        r = ops.linear_fwd(x, w, b, L.EPI_RESIDUAL, alpha=0.5, aux=aux)
        _close(r, aux.float() + 0.5 * (xf @ wf.t() + b), tol)
        dx = ops.linear_dgrad(dy, w)
        _close(dx, dyf @ wf, tol)
        dw = ops.linear_wgrad(dy, x, alpha=0.5)
        _close(dw, 0.5 * dyf.t() @ xf, tol if dtype == torch.bfloat16 else 1e-4)
        pre2 = _rand(M, K, dtype=dtype, seed=6)
        dpre = ops.linear_dgrad(dy, w, L.EPI_DSWISH, alpha=0.5, aux=pre2)
        s = torch.sigmoid(pre2.float())
        _close(dpre, 0.5 * (dyf @ wf) * (s * (1 + pre2.float() * (1 - s))), tol)
        cs = ops.colsum(dy, 0.5)
        _close(cs, 0.5 * dyf.sum(0), 1e-4 if dtype == torch.float32 else 1e-3)
    finally:
        ops.set_gemm_impl("auto")
        ops.set_tma(True)


CONVS = [  # N, H, W, Cin, Cout, k, stride
    (3, 8, 8, 64, 64, 3, 1), (37, 3, 3, 512, 512, 3, 1), (7, 6, 6, 256, 256, 3, 1), (3, 11, 11, 128, 128, 3, 1), (2, 22, 22, 64, 64, 3, 1), (2, 11, 11, 64, 128, 3, 2), (3, 6, 6, 128, 256, 3, 2), (5, 3, 3, 512, 512, 3, 1),
    (2, 11, 11, 64, 128, 1, 2), (4, 22, 22, 64, 64, 3, 1), (2, 22, 22, 64, 128, 3, 2), (5, 6, 6, 256, 512, 3, 2), (17, 6, 6, 256, 512, 3, 2),
    (3, 14, 14, 64, 64, 3, 1), (2, 9, 6, 64, 64, 3, 1), (150, 22, 22, 64, 64, 3, 1),   # halo-tile wgrad geometries ((W+2) % 8 == 0)
]


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 1e-5), ("tcgen05", torch.bfloat16, 1e-2), ("tcgen05-gather", torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("N,H,W,Ci,Co,k,s", CONVS)
def test_conv2d_fwd_dgrad_wgrad(impl, dtype, tol, N, H, W, Ci, Co, k, s):
    ops.set_tma(impl != "tcgen05-gather")
    ops.set_gemm_impl(impl.split("-")[0])
    try:
        x = _rand(N, H, W, Ci, dtype=dtype, seed=1)
        w = (_rand(Co, Ci, k, k, seed=2) / (Ci * k * k) ** 0.5).to(dtype)
        p = (k - 1) // 2
        g = ops.make_geom(N, 1, H, W, Ci, Co, (1, k, k), (1, s, s), (0, p, p))
        wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
        wd = w.permute(1, 2, 3, 0).reshape(Ci, -1).contiguous()
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        wr = w.float().requires_grad_(True)
        yr = F.conv2d(F.pad(xr, (p, k // 2, p, k // 2)), wr, None, stride=s)
        stats_r = torch.zeros(L.STATS_REPLICAS * 2 * Co, device=DEV)
        y = ops.conv_fwd(x, wp, g, colstats=stats_r)
        stats = stats_r.view(L.STATS_REPLICAS, 2 * Co).sum(0)
        yr_cl = yr.permute(0, 2, 3, 1).reshape(-1, Co)
        _close(y, yr_cl, tol)
        _close(stats[:Co], yr_cl.sum(0), 1e-2 if dtype == torch.bfloat16 else 1e-4)
        _close(stats[Co:], (yr_cl ** 2).sum(0), 1e-2 if dtype == torch.bfloat16 else 1e-4)
        dy = _rand(*y.shape, dtype=dtype, seed=3)
        yr.backward(dy.float().view(N, g.Ho, g.Wo, Co).permute(0, 3, 1, 2))
        dx = ops.conv_dgrad(dy, wd, g)
        _close(dx, xr.grad.permute(0, 2, 3, 1).reshape(-1, Ci), tol)
        dw = ops.conv_wgrad(dy, x, g)
        _close(dw, wr.grad.permute(0, 2, 3, 1).reshape(Co, -1), tol if dtype == torch.bfloat16 else 1e-4)
    finally:
        ops.set_gemm_impl("auto")
        ops.set_tma(True)


@pytest.mark.parametrize("N,H,W,Ci,Co", [(3, 22, 22, 64, 128), (4, 11, 11, 128, 256), (20, 6, 6, 256, 512), (2, 12, 10, 64, 64)])
def test_strided_dgrad_parity_classes_with_residual(N, H, W, Ci, Co):
    """stride-2 3x3 dgrad as four exact parity-class launches (tap table + scattered output rows), residual epilogue"""
    bf = torch.bfloat16
    g = ops.make_geom(N, 1, H, W, Ci, Co, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    w = (_rand(Co, Ci, 3, 3, seed=2) / (Ci * 9) ** 0.5).to(bf)
    wd = w.permute(1, 2, 3, 0).reshape(Ci, -1).contiguous()
    dy = _rand(N * g.Ho * g.Wo, Co, dtype=bf, seed=3)
    aux = _rand(N * H * W, Ci, dtype=bf, seed=4)
    xr = torch.zeros(N, Ci, H, W, device=DEV, requires_grad=True)
    yr = F.conv2d(F.pad(xr, (1, 1, 1, 1)), w.float(), None, stride=2)
    yr.backward(dy.float().view(N, g.Ho, g.Wo, Co).permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1).reshape(-1, Ci)
    _close(ops.conv_dgrad(dy, wd, g), want, 1e-2)
    _close(ops.conv_dgrad(dy, wd, g, epi=L.EPI_RESIDUAL, aux=aux), want + aux.float(), 1e-2)


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 1e-5), ("tcgen05", torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("B,T,H,W,Co,k,s", [(1, 4, 16, 16, 64, (5, 7, 7), (1, 2, 2)), (2, 1, 21, 80, 180, (1, 3, 3), (1, 2, 2)),
                                             (2, 5, 88, 88, 64, (5, 7, 7), (1, 2, 2))])
def test_single_channel_stem_conv_fwd_wgrad(impl, dtype, tol, B, T, H, W, Co, k, s):
    """C = 1 stems (video Conv3d 5x7x7, audio Conv2d 3x3): element-wise im2col gather inside the GEMM producer"""
    ops.set_gemm_impl(impl)
    try:
        taps = k[0] * k[1] * k[2]
        x = _rand(B, T, H, W, 1, dtype=dtype, seed=1)
        w = (_rand(Co, 1, *k, seed=2) / taps ** 0.5).to(dtype)
        b = _rand(Co, seed=3)
        pad = tuple((kk - 1) // 2 for kk in k)
        g = ops.make_geom(B, T, H, W, 1, Co, k, s, pad)
        y = ops.conv_fwd(x, w.reshape(Co, -1).contiguous(), g, bias=b)
        xr = x.float().view(B, 1, T, H, W)
        wr = w.float().requires_grad_(True)
        yr = F.conv3d(F.pad(xr, (pad[2], k[2] // 2, pad[1], k[1] // 2, pad[0], k[0] // 2)), wr, b, stride=s)
        _close(y, yr.permute(0, 2, 3, 4, 1).reshape(-1, Co), tol)
        dy = _rand(*y.shape, dtype=dtype, seed=4)
        yr.backward(dy.float().view(B, g.To, g.Ho, g.Wo, Co).permute(0, 4, 1, 2, 3))
        dw = ops.conv_wgrad(dy, x, g)
        _close(dw, wr.grad.reshape(Co, -1), tol if dtype == torch.bfloat16 else 1e-4)
    finally:
        ops.set_gemm_impl("auto")


@pytest.mark.parametrize("B,T,H,W", [(2, 5, 88, 88), (1, 3, 24, 88), (3, 1, 8, 88), (1, 7, 16, 88)])
def test_direct_stem3d_fwd_wgrad(B, T, H, W):
    """direct tcgen05 visual stem (im2col tile built in shared memory, no-swizzle UMMA operands) against conv3d in fp32 on the
    same bf16-rounded operands: output (+ bias), BatchNorm column sums from the epilogue, weight gradient"""
    bf = torch.bfloat16
    x = _rand(B, T, H, W, 1, dtype=bf, seed=1)
    w = (_rand(64, 1, 5, 7, 7, seed=2) / 245 ** 0.5).to(bf)
    b = _rand(64, seed=3)
    assert ops.stem3d_supported(x, 64, 5, 7, 7)
    stats = ops.gemm_stats_buffer(64, x.device)
    y = ops.stem3d_fwd(x, ops.stem3d_pack_weight(w.view(64, 1, 5, 7, 7)).contiguous(), b, colstats=stats)
    wr = w.float().requires_grad_(True)
    yr = F.conv3d(F.pad(x.float().view(B, 1, T, H, W), (3, 3, 3, 3, 2, 2)), wr, b, stride=(1, 2, 2))
    Ho, Wo = H // 2, W // 2
    yr2 = yr.permute(0, 2, 3, 4, 1).reshape(-1, 64)
    _close(y, yr2, 1e-2)
    st = stats.view(L.STATS_REPLICAS, 2, 64).sum(0)
    _close(st[0], yr2.sum(0), 2e-3)
    _close(st[1], (yr2 ** 2).sum(0), 2e-3)
    dy = _rand(*y.shape, dtype=bf, seed=4)
    yr.backward(dy.float().view(B, T, Ho, Wo, 64).permute(0, 4, 1, 2, 3))
    dw = ops.stem3d_wgrad(x, dy)
    _close(dw, wr.grad.reshape(64, 245), 1e-2)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("N,H,W,Co", [(2, 21, 80, 180), (3, 401, 80, 180), (1, 7, 10, 64)])
def test_audio_stem_conv2d_simt(dtype, tol, N, H, W, Co):
    """SIMT audio stem (Conv2d 1 -> Co, 3x3, stride 2, pad 1): output + bias, BatchNorm column sums, weight gradient"""
    x = _rand(N, H, W, dtype=dtype, seed=1)
    w = (_rand(Co, 9, seed=2) / 3).to(dtype)
    b = _rand(Co, seed=3)
    stats = ops.gemm_stats_buffer(Co, x.device)
    y = ops.stem2d_fwd(x, w, b, colstats=stats)
    wr = w.float().view(Co, 1, 3, 3).requires_grad_(True)
    yr = F.conv2d(F.pad(x.float().view(N, 1, H, W), (1, 1, 1, 1)), wr, b, stride=2)
    yr2 = yr.permute(0, 2, 3, 1).reshape(-1, Co)
    _close(y, yr2, tol)
    st = stats.view(L.STATS_REPLICAS, 2, Co).sum(0)
    _close(st[0], yr2.sum(0), 2e-3 if dtype == torch.bfloat16 else 1e-4)
    _close(st[1], (yr2 ** 2).sum(0), 2e-3 if dtype == torch.bfloat16 else 1e-4)
    dy = _rand(*y.shape, dtype=dtype, seed=4)
    yr.backward(dy.float().view(N, yr.shape[2], yr.shape[3], Co).permute(0, 3, 1, 2))
    _close(ops.stem2d_wgrad(x, dy), wr.grad.reshape(Co, 9), tol if dtype == torch.bfloat16 else 1e-4)


def test_im2col_single_channel():
    B, T, H, W = 2, 5, 20, 24
    for dtype in (torch.float32, torch.bfloat16):
        x = _rand(B, T, H, W, 1, dtype=dtype, seed=3)
        g = ops.make_geom(B, T, H, W, 1, 64, (5, 7, 7), (1, 2, 2), (2, 3, 3))
        col = ops.im2col_c1(x, g, 256)
        xr = F.pad(x.float().view(B, 1, T, H, W), (3, 3, 3, 3, 2, 2))
        ref = xr.unfold(2, 5, 1).unfold(3, 7, 2).unfold(4, 7, 2)          # (B,1,To,Ho,Wo,5,7,7)
        ref = ref.reshape(B * g.To * g.Ho * g.Wo, 245)
        assert torch.equal(col[:, :245].float(), ref)
        assert float(col[:, 245:].abs().max()) == 0.0
