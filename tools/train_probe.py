"""HBM roofline of the training-step kernels (csrc/train.cu) at BASELINE sizes, CUDA events, L2 flushed between launches
(a 512 MB memset), median of 20: fused Adam over the AV model's 61.7 M parameters, dropout (+ residual) on the stage-1
FFN hidden [12864 x 720] and output [12864 x 180] tensors, SpecAugment on [64 x 401 x 80], greedy decode on [64 x 51 x 256].
Also times the reference's path for the optimizer step on the same GPU (torch.optim.Adam foreach + clip_grad_norm_ + EMA loop).
usage: python tools/train_probe.py > gpurun_out/train_probe.log"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avec_b200
from avec_b200 import nnet, ops

dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm = None
for k, v in peaks.items():
    if "hbm" in k.lower() and isinstance(v, (int, float)):
        hbm = float(v) if hbm is None else max(hbm, float(v))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "GB/s": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / hbm, 3) if hbm else None}), flush=True)


print(json.dumps({"hbm_peak_GBps": hbm}))
# ---- fused Adam on the AV model
m = nnet.AudioVisualEfficientConformerInterCTC().to(dev)
n = sum(p.numel() for p in m.parameters())
opt = nnet.optimizers.Adam(m.parameters(), lr=nnet.schedulers.NoamDecayScheduler(10000, 360, 2), betas=(0.9, 0.98), eps=1e-9,
                           weight_decay=1e-6, grad_max_norm=None)
for p, gv in zip(m.parameters(), opt.grad_views()):
    gv.normal_(0, 1e-3)
    p.grad = gv
opt.step()
report("adam (p,g,m,v read; p,m,v written), 61.7 M parameters", timeit(opt.step), 28 * opt.flat()["n"])
opt2 = nnet.optimizers.Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6, grad_max_norm=5.0, ema_tau=0.999)
for p, gv in zip(m.parameters(), opt2.grad_views()):
    gv.normal_(0, 1e-3)
    p.grad = gv
opt2.step()
report("sumsq + adam + EMA (clip 5.0, tau 0.999)", timeit(opt2.step), (4 + 28 + 8) * opt2.flat()["n"])
# the reference's optimizer path on the same GPU: torch Adam + clip_grad_norm_ + per-tensor EMA loop (model.py:378-404)
ref_params = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters()]
ema = [p.detach().clone() for p in ref_params]
for p in ref_params:
    p.grad = torch.randn_like(p) * 1e-3
topt = torch.optim.Adam(ref_params, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)


def ref_step():
    torch.nn.utils.clip_grad_norm_(ref_params, 5.0)
    topt.step()
    for t, p in zip(ema, ref_params):
        t.mul_(0.999)
        t.add_((1 - 0.999) * p.detach())


ref_step()
report("reference path: torch.optim.Adam + clip_grad_norm_ + EMA loop (same GPU, eager)", timeit(ref_step, 5), 40 * n)
# ---- dropout
for rows, C, what in ((12864, 720, "FFN hidden"), (12864, 180, "FFN output + residual")):
    x = torch.randn(rows, C, device=dev).to(torch.bfloat16)
    res = torch.randn(rows, C, device=dev).to(torch.bfloat16) if "residual" in what else None
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.dropout(x, 0.1, 3, res=res, alpha=0.5, out=y))
    report(f"dropout {what} [{rows} x {C}] bf16", ms, rows * C * 2 * (3 if res is not None else 2))
# ---- SpecAugment / greedy decode
mel = torch.randn(64, 401, 80, device=dev)
ln = torch.full((64,), 401, device=dev)
report("spec_augment [64 x 401 x 80] fp32 (only masked elements are written)", timeit(lambda: ops.spec_augment_(mel, ln, 1)), 64 * 401 * 80 * 4 * 0.3)
logits = torch.randn(64, 51, 256, device=dev)
report("ctc_greedy_decode [64 x 51 x 256] fp32", timeit(lambda: ops.ctc_greedy_decode(logits, None)), 64 * 51 * 256 * 4)
