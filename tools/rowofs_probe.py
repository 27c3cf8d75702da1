"""Diagnostic: do UMMA shared-memory descriptors with SWIZZLE_128B accept start addresses that are 128-byte (one row)
but not 1024-byte aligned?  (needed to reuse one TMA-loaded halo tile for all 9 taps of a 3x3 convolution)
Run with AVEC_DEBUG_ROWOFS=r: the A tile is loaded r rows early and the descriptor starts r rows into the tile."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avec_b200 import ops

ofs = int(os.environ.get("AVEC_DEBUG_ROWOFS", "0"))
torch.manual_seed(0)
M, K, N = 1024, 256, 128
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / 16
y = ops.linear_fwd(x, w).float()
ref = x.float() @ w.float().t()
rows = torch.arange(M, device="cuda")
ok_rows = (rows % 128) < (128 - ofs)
err = (y - ref).abs().amax(dim=1)
print(f"rowofs={ofs}: max err on rows that stay inside the tile: {float(err[ok_rows].max()):.4f}  (scale {float(ref.abs().max()):.2f}); "
      f"rows expected garbage: {int((~ok_rows).sum())}, their max err {float(err[~ok_rows].max()) if ofs else 0:.3f}")
