"""diagnostic: library launches of one eager AV step (forward + 6 CTC losses + backward), three times in a row"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avec_b200
from avec_b200 import nnet, ops

dev = torch.device("cuda")
drop = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model = nnet.AudioVisualEfficientConformerInterCTC()
if drop == 0:
    nnet.zero_dropout(model)
model = model.to(dev).train()
ctc = nnet.CTCLoss(zero_infinity=True, assert_shorter=False)
video = torch.randn(B, 101, 88, 88, 1, device=dev).clamp(-1, 1)
audio = 0.1 * torch.randn(B, 64000, device=dev)
vl, al = torch.full((B,), 101, device=dev), torch.full((B,), 64000, device=dev)
labels, ll = torch.randint(1, 256, (B, 20), device=dev), torch.full((B,), 20, device=dev)
for i in range(3):
    ops.reset_launch_count()
    out = model((video, vl, audio, al))
    f = ops.launch_count()
    loss = sum(ctc((labels, ll), v) for v in out.values()) / len(out)
    l = ops.launch_count()
    loss.backward()
    torch.cuda.synchronize()
    print(f"step {i}: forward {f}, +losses {l}, +backward {ops.launch_count()}", flush=True)
