import torch
a = torch.load("gpurun_out/attn_mma.pt"); b = torch.load("gpurun_out/attn_simt.pt")
for k in a:
    for n, x, y in zip(("o", "probs", "dqkv", "de"), a[k], b[k]):
        err = (x - y).abs()
        print(k, n, "max abs", float(err.max()), "ref absmax", float(y.abs().max()), "rel", float((x - y).norm() / y.norm()), "nan", bool(torch.isnan(x).any()))
        if n == "dqkv" and float(err.max()) > 0.1 * float(y.abs().max()):
            B, T, H, d = k; D = H * d
            e3 = err.view(B, T, 3, H, d)
            print("   err by which(q,k,v):", e3.amax((0, 1, 3, 4)).tolist(), " by batch:", e3.amax((1, 2, 3, 4)).tolist(), " by head:", e3.amax((0, 1, 2, 4)).tolist())
            print("   err by token:", [round(v, 3) for v in e3.amax((0, 2, 3, 4)).tolist()])
