"""ResNet stage-1 BatchNorm backward (6464 x 22 x 22 x 64, ReLU, residual) for `ncu --set full -k regex:colreduce|bn_bwd_apply|bn_apply`."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops, _lib as L

dev, bf = "cuda", torch.bfloat16
rows, C = 6464 * 22 * 22, 64
u = torch.randn(rows, C, device=dev, dtype=bf)
dy = torch.randn(rows, C, device=dev, dtype=bf)
res = torch.randn(rows, C, device=dev, dtype=bf)
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
st = torch.stack([u.float().sum(0), (u.float() ** 2).sum(0)]).reshape(-1).contiguous()
buf = ops.bn_finalize(st, gam, bet, rows)
for _ in range(2):
    ops.bn_bwd(dy, u, buf, gam, L.ACT_RELU, res=res, want_dres=True)
    ops.bn_bwd(dy, u, buf, gam, L.ACT_RELU)
    ops.bn_apply(u, buf[0], buf[1], L.ACT_RELU, res=res)
torch.cuda.synchronize()
for name, fn in [("bn_bwd+res", lambda: ops.bn_bwd(dy, u, buf, gam, L.ACT_RELU, res=res, want_dres=True)), ("bn_bwd", lambda: ops.bn_bwd(dy, u, buf, gam, L.ACT_RELU)),
                 ("bn_apply+res", lambda: ops.bn_apply(u, buf[0], buf[1], L.ACT_RELU, res=res))]:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1):.3f} ms")
