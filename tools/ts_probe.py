"""Phase timeline of CTA (0,0,0) of representative tcgen05 GEMM launches (globaltimer, ns)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops, _lib as L

dev, bf = "cuda", torch.bfloat16
ts = torch.zeros(232, dtype=torch.int64, device=dev)
lib = L.load()
names = ["setup", "first k-block", "mainloop issue", "drain->accum", "epilogue", "dealloc"]


def run(label, fn):
    fn(); fn()
    torch.cuda.synchronize()
    ts.zero_()
    torch.cuda.synchronize()
    lib.avec_set_debug_timestamps(ts.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    lib.avec_set_debug_timestamps(0)
    t = ts.cpu().tolist()
    d = [t[i + 1] - t[i] for i in range(6)]
    if t[8 + 5 * 6]:
        for j in range(4, 9):
            b = 8 + 5 * j
            if not t[b + 2]:
                break
            print(f"   tile {j}: kb0 ready +{t[b + 3] - t[b - 5 + 2]} (rel. prev epilogue end), last MMA issued +{t[b + 4] - t[b + 3]}, "
                  f"epilogue waited {t[b + 1] - t[b]}, epilogue ran {t[b + 2] - t[b + 1]}, period {t[b + 2] - t[b - 5 + 2]}")
    if t[168]:
        m = [t[168 + i] - t[168] for i in range(32) if t[168 + i]]
        q = [t[200 + i] - t[168] for i in range(32) if t[200 + i]]
        print("   tile 6 MMA thread (cycles rel. first full-wait done; pairs = wait done, issued):", m)
        print("   tile 6 producer   (same origin; pairs = empty-wait done, TMA issued):", q)
    print(f"{label}: kernel {e0.elapsed_time(e1) * 1000:.1f} us; CTA0 phases (ns): " + ", ".join(f"{n} {v}" for n, v in zip(names, d)) + f"; total {t[6] - t[0]}")


x, w, b = torch.randn(6464, 256, device=dev, dtype=bf), torch.randn(256, 256, device=dev, dtype=bf), torch.randn(256, device=dev)
aux = torch.randn(6464, 256, device=dev, dtype=bf)
run("small linear 6464x256x256 residual", lambda: ops.linear_fwd(x, w, b, L.EPI_RESIDUAL, aux=aux))
x2, w2 = torch.randn(6464, 1024, device=dev, dtype=bf), torch.randn(256, 1024, device=dev, dtype=bf)
run("linear 6464x1024->256", lambda: ops.linear_fwd(x2, w2, b))
N, H, W, C = 6464, 22, 22, 64
xi = torch.randn(N, H, W, C, device=dev, dtype=bf)
g = ops.make_geom(N, 1, H, W, C, C, (1, 3, 3), (1, 1, 1), (0, 1, 1))
wp = torch.randn(C, 9 * C, device=dev, dtype=bf)
st = torch.zeros(32 * 2 * C, device=dev)
run("conv stage-1 fwd + stats", lambda: ops.conv_fwd(xi, wp, g, colstats=st))
run("conv stage-1 fwd no stats", lambda: ops.conv_fwd(xi, wp, g))
N, H, W, C = 6464, 6, 6, 256
xi3 = torch.randn(N, H, W, C, device=dev, dtype=bf)
g3 = ops.make_geom(N, 1, H, W, C, C, (1, 3, 3), (1, 1, 1), (0, 1, 1))
wp3 = torch.randn(C, 9 * C, device=dev, dtype=bf)
run("conv stage-3 fwd", lambda: ops.conv_fwd(xi3, wp3, g3))
xb, wb = torch.randn(8192, 8192, device=dev, dtype=bf), torch.randn(8192, 8192, device=dev, dtype=bf)
run("8192^3", lambda: ops.linear_fwd(xb, wb))
