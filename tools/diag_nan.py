import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import avec_b200, seeded
from avec_b200 import nnet, ops
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_dropin import _batch
DEV = "cuda"
inputs, targets = _batch()
inputs = [x.to(DEV) for x in inputs]; targets = tuple(x.to(DEV) for x in targets)
avec_b200.set_compute_dtype(torch.bfloat16)
torch.manual_seed(0)
m = nnet.AudioVisualEfficientConformerInterCTC().to(DEV).train()
m.compile(losses=nnet.CTCLoss(zero_infinity=True, assert_shorter=False), optimizer="Adam")
for step in range(3):
    out = m(tuple(inputs))
    loss = m.compute_loss(out, targets)
    for p in m.parameters():
        p.grad = None
    loss.backward()
    bad = [k for k, p in m.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    print("step", step, "loss", float(loss), "bad grads:", len(bad), bad[:8], flush=True)
    m.optimizer.step()
    badw = [k for k, p in m.named_parameters() if not torch.isfinite(p).all()]
    print("   bad weights:", len(badw), badw[:8], "info", m.optimizer.last_info(), flush=True)
    if bad or badw:
        break
