"""Does alternating the 227 KB-shared-memory tcgen05 kernel with small elementwise kernels cost SM reconfiguration time?
Times CUDA graphs of (a) 100 small GEMMs, (b) 100 LayerNorms, (c) 100 alternating pairs."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops, _lib as L

dev, bf = "cuda", torch.bfloat16
M, D = 6464, 256
x = torch.randn(M, D, device=dev, dtype=bf)
w = torch.randn(D, D, device=dev, dtype=bf)
b = torch.randn(D, device=dev)
g, be = torch.ones(D, device=dev), torch.zeros(D, device=dev)
x3 = x.view(64, 101, D)


def gemm():
    return ops.linear_fwd(x, w, b)


def ln():
    return ops.layernorm_fwd(x3, g, be)


def colsum():
    return ops.colsum(x)


def graph_time(fns, reps=100):
    for f in fns:
        f()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            for f in fns:
                f()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / reps * 1000.0   # us per repetition


ta, tb, tc_ = graph_time([gemm]), graph_time([ln]), graph_time([colsum])
tab, tac = graph_time([gemm, ln]), graph_time([gemm, colsum])
tabc = graph_time([gemm, ln, gemm, colsum])
print(f"gemm {ta:.2f} us, layernorm {tb:.2f} us, colsum {tc_:.2f} us")
print(f"gemm+ln pair {tab:.2f} us (sum of parts {ta + tb:.2f}); gemm+colsum pair {tac:.2f} us (sum {ta + tc_:.2f})")
print(f"gemm, ln, gemm, colsum {tabc:.2f} us (sum {2 * ta + tb + tc_:.2f})")
