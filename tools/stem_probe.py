"""Time the direct visual-stem kernels (and the im2col + GEMM path they replace) at B = 64 x 101 frames of 88 x 88."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops

dev, bf = "cuda", torch.bfloat16
B, T = int(os.environ.get("B", "64")), 101
x = torch.randn(B, T, 88, 88, 1, device=dev, dtype=bf)
w = (torch.randn(64, 1, 5, 7, 7, device=dev) / 16).to(bf)
wp = ops.stem3d_pack_weight(w).contiguous()
b = torch.randn(64, device=dev)
sites = B * T * 44 * 44
dy = torch.randn(sites, 64, device=dev, dtype=bf)
st = ops.gemm_stats_buffer(64, x.device)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


fl = 2.0 * sites * 64 * 245
for name, fn in [("stem3d_fwd", lambda: ops.stem3d_fwd(x, wp, b, colstats=st)), ("stem3d_fwd_nostats", lambda: ops.stem3d_fwd(x, wp, b)),
                 ("stem3d_wgrad", lambda: ops.stem3d_wgrad(x, dy))]:
    ms = timeit(fn)
    print(f"dbg={os.environ.get('AVEC_STEM_DBG', '0')} {name}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
