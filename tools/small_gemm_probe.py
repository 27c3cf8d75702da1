"""Conformer-sized GEMMs (few k-blocks, ~50 row tiles) under different tile widths: AVEC_MAX_BN=256|128|64 python tools/small_gemm_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops, _lib as L
dev, bf = "cuda", torch.bfloat16


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0


print("AVEC_MAX_BN", os.environ.get("AVEC_MAX_BN"))
for M, K, N in [(6464, 256, 1024), (6464, 1024, 256), (6464, 256, 256), (6464, 256, 768), (3264, 360, 1440), (3264, 1440, 360), (12864, 180, 720), (12864, 720, 180)]:
    x, w, b = torch.randn(M, K, device=dev, dtype=bf), ops.convert(torch.randn(N, K, device=dev), bf, pad=True), torch.randn(N, device=dev)
    dy, pre, res = torch.randn(M, N, device=dev, dtype=bf), torch.randn(M, N, device=dev, dtype=bf), torch.randn(M, N, device=dev, dtype=bf)
    rows = [("fwd linear", lambda: ops.linear_fwd(x, w, b)), ("fwd swish+pre", lambda: ops.linear_fwd(x, w, b, L.EPI_SWISH, want_pre=True)),
            ("fwd residual", lambda: ops.linear_fwd(x, w, b, L.EPI_RESIDUAL, alpha=0.5, aux=res)),
            ("dgrad", lambda: ops.linear_dgrad(dy, w)), ("wgrad", lambda: ops.linear_wgrad(dy, x))]
    print(f"M{M} K{K} N{N}: " + "  ".join(f"{n} {timeit(f):.1f}us" for n, f in rows), flush=True)
