"""Time individual avec_gemm launches (CUDA events, L2 flushed between repetitions) at the BASELINE shapes."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops, _lib as L

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


res = []
bf = torch.bfloat16
# plain linears (M, K, N)
for M, K, N in [(12864, 180, 720), (12864, 720, 180), (12864, 180, 540), (6464, 256, 1024), (6464, 1024, 256), (3264, 360, 1440),
                (3264, 1440, 360), (12864, 7200, 180), (8192, 8192, 8192)]:
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf)
    b = torch.randn(N, device=dev)
    dy = torch.randn(M, N, device=dev, dtype=bf)
    for name, fn in [("fwd", lambda: ops.linear_fwd(x, w, b, L.EPI_SWISH, want_pre=True)), ("dgrad", lambda: ops.linear_dgrad(dy, w)),
                     ("wgrad", lambda: ops.linear_wgrad(dy, x))]:
        ms = timeit(fn)
        res.append({"op": f"linear_{name}", "M": M, "K": K, "N": N, "ms": ms, "tflops": 2 * M * K * N / ms / 1e9})
        print(res[-1], flush=True)
# ResNet convs (N images, H, W, Cin, Cout, k, stride) at B=64 x 101 frames
for N, H, W, Ci, Co, k, s in [(6464, 22, 22, 64, 64, 3, 1), (6464, 22, 22, 64, 128, 3, 2), (6464, 11, 11, 128, 128, 3, 1),
                              (6464, 11, 11, 128, 256, 3, 2), (6464, 6, 6, 256, 256, 3, 1), (6464, 6, 6, 256, 512, 3, 2), (6464, 3, 3, 512, 512, 3, 1)]:
    x = torch.randn(N, H, W, Ci, device=dev, dtype=bf)
    p = (k - 1) // 2
    g = ops.make_geom(N, 1, H, W, Ci, Co, (1, k, k), (1, s, s), (0, p, p))
    wp = torch.randn(Co, k * k * Ci, device=dev, dtype=bf)
    wd = torch.randn(Ci, k * k * Co, device=dev, dtype=bf)
    dy = torch.randn(ops.geom_sites(g), Co, device=dev, dtype=bf)
    st = torch.zeros(32 * 2 * Co, device=dev)
    fl = 2.0 * ops.geom_sites(g) * Co * k * k * Ci
    for name, fn in [("fwd", lambda: ops.conv_fwd(x, wp, g, colstats=st)), ("dgrad", lambda: ops.conv_dgrad(dy, wd, g)),
                     ("wgrad", lambda: ops.conv_wgrad(dy, x, g))]:
        ms = timeit(fn)
        res.append({"op": f"conv_{name}", "N": N, "HW": H, "Ci": Ci, "Co": Co, "s": s, "ms": ms, "tflops": fl / ms / 1e9})
        print(res[-1], flush=True)
# video stem
B, T = 64, 101
x = torch.randn(B, T, 88, 88, 1, device=dev, dtype=bf)
g = ops.make_geom(B, T, 88, 88, 1, 64, (5, 7, 7), (1, 2, 2), (2, 3, 3))
wp = torch.randn(64, 245, device=dev, dtype=bf)
dy = torch.randn(ops.geom_sites(g), 64, device=dev, dtype=bf)
fl = 2.0 * ops.geom_sites(g) * 64 * 245
col = ops.im2col_c1(x, g, 256)
wp256 = torch.randn(64, 256, device=dev, dtype=bf)
st = torch.zeros(32 * 128, device=dev)
for name, fn in [("im2col", lambda: ops.im2col_c1(x, g, 256)), ("fwd", lambda: ops.linear_fwd(col, wp256, None, colstats=st)),
                 ("wgrad", lambda: ops.linear_wgrad(dy, col))]:
    ms = timeit(fn, 3)
    res.append({"op": f"stem3d_{name}", "ms": ms, "tflops": fl / ms / 1e9})
    print(res[-1], flush=True)
json.dump(res, open("gpurun_out/gemm_probe.json", "w"), indent=1)
