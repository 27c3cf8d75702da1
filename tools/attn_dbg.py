import os, sys, torch, subprocess
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import seeded
from avec_b200 import ops
bf = torch.bfloat16
DEV = "cuda"
def run():
    res = {}
    for (B, T, H, d) in [(3, 9, 4, 90), (3, 7, 4, 45), (3, 6, 4, 64), (3, 17, 4, 64), (2, 51, 4, 90), (2, 101, 4, 64)]:
        D = H * d
        qkv = seeded.randn("qkv", (B * T, 3 * D), 5, 0.5).to(DEV).to(bf)
        e = seeded.randn("e", (2 * T - 1, D), 5, 0.5).to(DEV).to(bf)
        klen = torch.tensor([T] + [max(1, T - 3 - 2 * i) for i in range(B - 1)], device=DEV, dtype=torch.int32)
        o, probs = ops.relpos_attn_fwd(qkv, e, klen, T, B, T, H, d)
        do = seeded.randn("do", (B * T, D), 5).to(DEV).to(bf)
        dqkv, de, _, _ = ops.relpos_attn_bwd(do, qkv, e, probs, B, T, H, d)
        torch.cuda.synchronize()
        res[(B, T, H, d)] = [t.float().cpu() for t in (o, probs, dqkv, de)]
    return res
r = run()
torch.save(r, sys.argv[1])
