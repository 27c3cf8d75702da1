"""BASELINE.json configs[4]: patch vs grouped vs regular attention of the audio-only Efficient Conformer, sequence-length
sweep 100 -> 1600 mel frames (1 s .. 16 s of audio), forward + 4 CTC losses + backward on one B200 (training graph: dropout 0.1 + SpecAugment; the
step is captured in a CUDA graph, 3 warm-up + 5 timed replays, CUDA events, inputs resident in HBM).  Prints one JSON line per (attention, frames) and a markdown table.
usage: python tools/ablation_sweep.py [batch] > gpurun_out/ablation.log"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avec_b200
from avec_b200 import nnet

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
FORCE_LONG = os.environ.get("AVEC_ATTN_LONG") == "1"      # diagnostics: key-tiled attention kernels for every length
if FORCE_LONG:
    from avec_b200 import _lib
    _lib.load().avec_set_attention_long(1)
FRAMES = [int(f) for f in os.environ.get("FRAMES", "100,200,400,800,1600").split(",")]
avec_b200.set_compute_dtype(torch.bfloat16)
ctc = nnet.CTCLoss(zero_infinity=True, assert_shorter=False)
rows = []
for att in ("patch", "grouped", "regular"):  # noqa: C901
    torch.manual_seed(0)
    model = nnet.AudioEfficientConformerInterCTC(att_type=att).to(dev).train()
    for frames in FRAMES:
        L = (frames - 1) * 160
        audio = 0.1 * torch.randn(B, L, device=dev)
        alen = torch.full((B,), L, device=dev)
        labels, ll = torch.randint(1, 256, (B, 10), device=dev), torch.full((B,), 10, device=dev)

        def step():
            for p in model.parameters():
                p.grad = None
            out = model((audio, alen))
            loss = sum(ctc((labels, ll), v) for v in out.values()) / len(out)
            loss.backward()
            return loss

        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    loss = step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = step()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            rec = {"att": att, "frames": frames, "batch": B, "ms_per_step": round(ms, 3), "utt_per_s": round(B / ms * 1e3, 1),
                   "frames_per_s": round(B * frames / ms * 1e3), "loss_finite": bool(torch.isfinite(loss)),
                   "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
        except Exception as e:  # noqa: BLE001
            rec = {"att": att, "frames": frames, "batch": B, "error": f"{type(e).__name__}: {str(e)[:160]}"}
        graph = None
        torch.cuda.reset_peak_memory_stats()
        rows.append(rec)
        print(json.dumps(rec), flush=True)
print("\n| attention | frames | ms / step | utterances/s | mel frames/s | peak GB |\n|---|---:|---:|---:|---:|---:|")
for r in rows:
    if "error" in r:
        print(f"| {r['att']} | {r['frames']} | {r['error']} | | | |")
    else:
        print(f"| {r['att']} | {r['frames']} | {r['ms_per_step']} | {r['utt_per_s']} | {r['frames_per_s']} | {r['peak_mem_gb']} |")
