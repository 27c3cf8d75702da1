"""diagnostic: per-head relative error of the AV / AO / VO models at the benchmark shape against the fp32 restatement (GPU)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import avec_b200, seeded
from avec_b200 import nnet
from oracle import restate
from common import rel_err
DEV = "cuda"
B, Ls = int(os.environ.get("B", 8)), 64000
g = torch.Generator().manual_seed(1234)
audio = (0.1 * torch.randn(B, Ls, generator=g)).to(DEV)
ragged = int(os.environ.get("RAGGED", 1))
alen = torch.tensor([Ls - (1280 * (i % 4) if ragged else 0) for i in range(B)], device=DEV)
for kind in os.environ.get("KINDS", "AO,VO,AV").split(","):
    Tv = 101 if kind == "AV" else 100
    video = torch.randn(B, Tv, 88, 88, 1, generator=g).clamp_(-1, 1).to(DEV)
    vlen = (alen // 640 + 1).clamp(max=Tv)
    for dt in (torch.float32, torch.bfloat16):
        avec_b200.set_compute_dtype(dt)
        cls = {"AV": nnet.AudioVisualEfficientConformerInterCTC, "AO": nnet.AudioEfficientConformerInterCTC, "VO": nnet.VisualEfficientConformerInterCTC}[kind]
        m = cls()
        m.load_state_dict(seeded.seeded_state_dict(m, 11))
        nnet.zero_dropout(m)
        m = m.to(DEV).train()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        inp = {"AV": (video, vlen, audio, alen), "AO": (audio, alen), "VO": (video, vlen)}[kind]
        with torch.no_grad():
            out = m(inp)
            want = {"AV": lambda: restate.av_model(sd, video, vlen, audio, alen), "AO": lambda: restate.ao_model(sd, audio, alen),
                    "VO": lambda: restate.vo_model(sd, video, vlen)}[kind]()
        print(kind, dt, {k: round(rel_err(out[k][0], want[k][0]), 5) for k in want}, flush=True)
