"""Exactly two eager AV training steps (B = 64) for `ncu --metrics gpu__time_duration.sum`: the first warms up, the second is
the one tools/summarize_launches.py condenses (launches between the last two stft_mel_log kernels + a trailing marker step).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avec_b200
from avec_b200 import nnet
import bench

dev = torch.device("cuda", 0)
avec_b200.set_compute_dtype(torch.bfloat16)
torch.manual_seed(1234)
model = nnet.AudioVisualEfficientConformerInterCTC()          # DROPOUT=0: the deterministic parity graph; default: the training graph
if os.environ.get("DROPOUT", "0.1") == "0":
    nnet.zero_dropout(model)
model = model.to(dev).train()
ctc = nnet.CTCLoss(zero_infinity=True, assert_shorter=False)
d = bench.synth_inputs("AV", int(os.environ.get("B", "64")), dev)
for step in range(3):   # the third step only contributes its first (video + stft) kernels as the closing marker
    for p in model.parameters():
        p.grad = None
    avec_b200.invalidate_weights()     # as bench.py: the bf16 weight copies are rebuilt every step
    out = model(bench.model_inputs("AV", d))
    if step == 2:
        break
    loss = sum(ctc((d["labels"], d["llen"]), v) for v in out.values()) / len(out)
    loss.backward()
torch.cuda.synchronize()
