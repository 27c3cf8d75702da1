#!/usr/bin/env python
"""bench.py - utterances/s of the Audio-Visual Efficient Conformer (AVEC) encoder forward+backward on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...              # the reference's own CPU implementation on the host cores

Workload (BASELINE.json metric / configs[3]): AV EffConfInterCTC, per-GPU batch 64, 4 s of 16 kHz audio (64000
samples) + 101 frames of 88x88 video (Tv = Ta // 640 + 1), synthetic data, random-init weights,
train-mode forward (batch-stat BatchNorm; dropout 0.1 + SpecAugment as in the reference's training graph, --dropout 0 = the
deterministic parity configuration) + CTC losses on the 6 heads + backward; the bf16 kernel-layout weight copies are rebuilt
from the fp32 masters inside every step.
One step = one batch.  `value` times steps with inputs resident in HBM; `e2e` times the public API call
model((video, vlen, audio, alen)) fed from pinned host memory, H2D copies and the D2H read of the loss inside the timed
region.  Under torchrun (N > 1) the gradients are all-reduced over NCCL inside the captured step (pure data parallel, local BN).
Beside the number (rank 0): `roofline` (algorithmic FLOPs of the step / CUDA-event time inside the tcgen05 kernels, per family),
`parity` (the timed model and batch through the dropout-free graph against the fp32 oracle on the same GPU), `incumbent` (the
reference's graph as eager bf16-autocast PyTorch on the same GPU: the unmodified reference when its tree is present) and
`cpu_baseline` (the reference's CPU path on a bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)



def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="avec_b200", choices=["avec_b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="AV", choices=["AV", "AO", "VO"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-baseline", type=int, default=1, help="time the CPU restatement on a bounded sample (rank 0, N=1)")
    ap.add_argument("--cpu-sample", type=int, default=8, help="utterances in the CPU baseline sample")
    ap.add_argument("--loss", default="ctc", choices=["ctc", "sum"])
    ap.add_argument("--optimizer", default="none", choices=["none", "adam"], help="adam: add the fused optimizer step (Adam + Noam LR + "
                    "global-norm clip, avec_b200.nnet.optimizers.Adam) to every step; the headline metric is forward + backward (none)")
    ap.add_argument("--overlap", type=int, default=1, help="AV model: audio encoder on a second stream, concurrent with the video encoder")
    ap.add_argument("--dropout", type=float, default=0.1, help="0.1 = the reference's training graph (dropout at every site + "
                    "SpecAugment, networks.py:327,347-353); 0 = the deterministic parity graph (dropout off, SpecAugment bypassed)")
    ap.add_argument("--graph", type=int, default=1, help="capture forward+backward in one CUDA graph (falls back to eager if capture fails)")
    ap.add_argument("--sync-bn", type=int, default=0, help="N > 1: the reference's SyncBatchNorm default (model.py:59-61) instead of local statistics")
    ap.add_argument("--bucket-mb", type=int, default=4096, help="N > 1: gradient all-reduce bucket size.  Default = ONE bucket, reduced right after the "
                    "backward inside the captured graph (2 x B200: 47.1 ms / step); 32 MB buckets overlapped with the backward measured SLOWER "
                    "(50.3 ms): the NCCL kernels take SMs away from the persistent one-CTA-per-SM conv / GEMM kernels (profiles/r02_multigpu.md)")
    ap.add_argument("--parity-check", type=int, default=1, help="compare the benchmarked model / batch with the fp32 oracle on the same GPU (rank 0)")
    ap.add_argument("--incumbent", type=int, default=1, help="time the reference graph as eager bf16-autocast PyTorch on the same GPU (N=1)")
    ap.add_argument("--shape-table", default="", help="write the per-shape tcgen05 launch table (ms, TFLOP/s) to this JSON file")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------- helpers
def synth_inputs(model, B, device, seed=1234, pinned=False):
    g = torch.Generator().manual_seed(seed)
    Ls, Tv = 64000, (101 if model == "AV" else 100)
    kw = dict(pin_memory=pinned)
    audio = (0.1 * torch.randn(B, Ls, generator=g)).contiguous()
    video = torch.randn(B, Tv, 88, 88, 1, generator=g).clamp_(-1, 1).contiguous()
    alen = torch.full((B,), Ls, dtype=torch.long)
    vlen = torch.full((B,), Tv, dtype=torch.long)
    labels = torch.randint(1, 256, (B, 20), generator=g)
    llen = torch.full((B,), 20, dtype=torch.long)
    if pinned:
        audio, video = audio.pin_memory(), video.pin_memory()
    host = dict(audio=audio, video=video, alen=alen, vlen=vlen, labels=labels, llen=llen)
    if device is None:
        return host
    return {k: v.to(device) for k, v in host.items()}


def model_inputs(model, d):
    if model == "AV":
        return (d["video"], d["vlen"], d["audio"], d["alen"])
    if model == "AO":
        return (d["audio"], d["alen"])
    return (d["video"], d["vlen"])


class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def gemm_traffic():
    """DRAM bytes (read + write) of all tcgen05 GEMM / conv launches of one step, from the committed ncu pass
    (profiles/r02_gemm_traffic.json: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over tools/one_step.py, tools/gemm_traffic.py)"""
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p)).get("dram_bytes_per_step")
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"sustained": d.get("bf16_tflops_sustained", 1415.1), "burst": d.get("bf16_tflops", 1643.8), "hbm": d.get("hbm_gbs", 6449.1), "which": "measured"}
    return {"sustained": 1400.0, "burst": 1600.0, "hbm": 6650.0, "which": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------- baselines: reference / port, CPU / GPU
GFLOP_PER_UTT = {"AV": 216.38, "AO": 20.24, "VO": 198.24}   # SURVEY section 8d (torch flop counter on the reference, fwd+bwd)


def _baseline_callable(args, B, device, state_dict=None):
    """forward + CTC losses + backward of the named model as a zero-argument callable, built from the UNMODIFIED reference when a
    reference tree is present (/root/reference here, baseline/_ref on the GPU box: kind "reference", the reference's zoo model in
    train() - dropout 0.1, SpecAugment, its own CTCLoss), else from the pinned plain-torch restatement oracle/restate.py (kind
    "port": same graph, dropout / SpecAugment off)."""
    from oracle import ref_import, restate
    import torch.nn.functional as F
    d = synth_inputs(args.model, B, None)
    d = {k: v.to(device) for k, v in d.items()}
    if ref_import.available() and not os.environ.get("AVEC_BENCH_FORCE_PORT"):
        ref = ref_import.import_reference()
        if getattr(ref, "__avec_b200_patched__", False):
            raise RuntimeError("the reference arm needs the UNPATCHED reference")
        cls = {"AV": ref.AudioVisualEfficientConformerInterCTC, "AO": ref.AudioEfficientConformerInterCTC,
               "VO": ref.VisualEfficientConformerInterCTC}[args.model]
        torch.manual_seed(1234)
        m = cls()
        if state_dict is not None:
            m.load_state_dict(state_dict)
        m = m.to(device).train()
        ctc = ref.CTCLoss(zero_infinity=True, assert_shorter=False)
        params = [p for p in m.parameters()]

        def step():
            for p in params:
                p.grad = None
            out = m.forward(list(model_inputs(args.model, d)))
            loss = sum(ctc((d["labels"], d["llen"]), v) for v in out.values()) / len(out)
            loss.backward()
            return loss.detach()
        return step, "reference", "the unmodified reference's zoo model (train(): dropout 0.1 + SpecAugment) + its CTCLoss on every head"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from avec_b200 import nnet
    cls = {"AV": nnet.AudioVisualEfficientConformerInterCTC, "AO": nnet.AudioEfficientConformerInterCTC, "VO": nnet.VisualEfficientConformerInterCTC}[args.model]
    if state_dict is None:
        torch.manual_seed(1234)
        state_dict = cls().state_dict()
    sd = {k: v.detach().clone().to(device).requires_grad_(v.is_floating_point() and "running" not in k and "Spectrogram" not in k and "MelScale" not in k)
          for k, v in state_dict.items()}
    fn = {"AV": lambda: restate.av_model(sd, d["video"], d["vlen"], d["audio"], d["alen"]),
          "AO": lambda: restate.ao_model(sd, d["audio"], d["alen"]),
          "VO": lambda: restate.vo_model(sd, d["video"], d["vlen"])}[args.model]

    def ctc_of(v):
        logp = F.log_softmax(v[0].float(), dim=-1).transpose(0, 1)
        return F.ctc_loss(logp, d["labels"], v[1].to(torch.long), d["llen"], blank=0, reduction="none", zero_infinity=True).mean()

    def step():
        for v in sd.values():
            v.grad = None
        out = fn()
        loss = sum(ctc_of(v) for v in out.values()) / len(out)
        loss.backward()
        return loss.detach()
    return step, "port", "the pinned plain-torch restatement oracle/restate.py (dropout / SpecAugment off) + F.ctc_loss on every head"


def cpu_baseline(args, steps=2, warmup=0, sample=None):
    """the reference's CPU implementation of the path on the box's host cores, on a bounded sample of the workload"""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = sample or args.cpu_sample
    step, kind, what = _baseline_callable(args, B, torch.device("cpu"))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    sec = (time.perf_counter() - t0) / steps
    return {"value": B / sec, "unit": "utterances/s", "cores": cores, "kind": kind,
            "sample": f"{B} utterances per step (4 s audio{'' if args.model == 'AO' else ' + video'}), {warmup} warm-up + {steps} timed fwd+CTC+bwd steps of the {args.model} "
                      f"model in fp32 on {cores} host threads: {what}", "seconds": sec, "steps_run": steps, "warmup_run": warmup}


def gpu_incumbent(args, dev, state_dict, steps=3, warmup=1):
    """the incumbent on the SAME B200: the reference's graph (live reference when a tree is present, else the
    restatement) as eager PyTorch under bf16 autocast - cuBLAS / cuDNN / ATen kernels, no code of this repository on the path"""
    B = args.batch
    step, kind, what = _baseline_callable(args, B, dev, state_dict)

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return step()
    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / (ms / 1000.0), "unit": "utterances/s", "ms_per_step": ms, "kind": kind, "dtype": "bf16 autocast (eager PyTorch: cuBLAS / cuDNN / ATen)",
            "what": f"{what}; per-GPU batch {B}, {warmup} warm-up + {steps} timed steps, inputs resident"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=max(1, args.steps), warmup=max(0, args.warmup), sample=args.cpu_sample)
    line = {"impl": "reference", "metric": "utterances/sec fwd+bwd (4s audio+video)", "value": cb["value"], "unit": "utterances/s",
            "n_gpus": args.gpus, "steps": cb["steps_run"], "warmup": cb["warmup_run"], "ms_per_step": 1000.0 * cb["seconds"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} EffConfInterCTC fwd+bwd (6 CTC heads), bounded sample of {args.cpu_sample} utterances per step "
                                   f"(4 s audio + video) on the host cores"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def parity_check(args, model, resident, dev):
    """the benchmarked configuration against the oracle: the SAME model, weights and full-size batch through
    the deterministic parity graph (dropout 0, SpecAugment off) in the benchmark's precision, next to the fp32 restatement
    (oracle/restate.py, pinned to the unmodified reference's fixtures) run on the same GPU.  Total CTC loss, logits and greedy
    alignment agreement; raises if the loss leaves the bf16 envelope."""
    from oracle import restate
    from avec_b200 import nnet
    import torch.nn.functional as F
    saved = [(m, m.p) for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    specs = [m for m in model.modules() if isinstance(m, nnet.SpecAugment)]
    spec_on = [m.enabled for m in specs]
    buffers = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
    try:
        nnet.zero_dropout(model)
        with torch.no_grad():
            out = model(model_inputs(args.model, resident))
        ctc = nnet.CTCLoss(zero_infinity=True, assert_shorter=False)
        with torch.no_grad():
            loss = float(sum(ctc((resident["labels"], resident["llen"]), v) for v in out.values()) / len(out))
        sd = {k: v.detach().clone().float() for k, v in model.state_dict().items()}
        for k, v in buffers.items():
            sd[k] = v.clone()
        with torch.no_grad():
            want = {"AV": lambda: restate.av_model(sd, resident["video"], resident["vlen"], resident["audio"], resident["alen"]),
                    "AO": lambda: restate.ao_model(sd, resident["audio"], resident["alen"]),
                    "VO": lambda: restate.vo_model(sd, resident["video"], resident["vlen"])}[args.model]()

            def ctc_of(v):
                logp = F.log_softmax(v[0].float(), dim=-1).transpose(0, 1)
                return F.ctc_loss(logp, resident["labels"], v[1].to(torch.long), resident["llen"], blank=0, reduction="none", zero_infinity=True).mean()
            oracle_loss = float(sum(ctc_of(v) for v in want.values()) / len(want))
        a, b = out["outputs"][0].float(), want["outputs"][0].float()
        rel = float((a - b).norm() / b.norm())
        agree = float((a.argmax(-1) == b.argmax(-1)).float().mean())
        top2 = b.topk(2, dim=-1).values
        clear = (top2[..., 0] - top2[..., 1]) > 0.05
        agree_clear = float((a.argmax(-1) == b.argmax(-1))[clear].float().mean()) if bool(clear.any()) else 1.0
    finally:
        for m, p in saved:
            m.p = p
        for m, e in zip(specs, spec_on):
            m.enabled = e
        model.load_state_dict(buffers, strict=False)
    res = {"graph": "parity (dropout 0, SpecAugment off), same weights / batch as the timed steps", "ctc_loss": loss, "oracle_ctc_loss": oracle_loss,
           "loss_rel_err": abs(loss - oracle_loss) / abs(oracle_loss), "logits_rel_l2": rel, "greedy_agreement": agree,
           "greedy_agreement_margin_gt_0.05": agree_clear, "oracle": "oracle/restate.py fp32 on the same GPU"}
    tol = 3e-2 if args.dtype == "bf16" else 1e-3
    if not (res["loss_rel_err"] <= tol):
        raise AssertionError(f"bench parity check failed: CTC loss {loss} vs oracle {oracle_loss} (rel {res['loss_rel_err']:.3e} > {tol})")
    return res


# ------------------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import avec_b200
    from avec_b200 import nnet, ops, parallel
    from avec_b200 import functional as AF
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    avec_b200.set_compute_dtype(dtype)

    torch.manual_seed(1234 + rank)
    cls = {"AV": nnet.AudioVisualEfficientConformerInterCTC, "AO": nnet.AudioEfficientConformerInterCTC, "VO": nnet.VisualEfficientConformerInterCTC}[args.model]
    model = cls()
    if args.dropout == 0:
        nnet.zero_dropout(model)
    else:
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = args.dropout
    model = model.to(dev).train()
    if args.model == "AV":
        model.encoder.overlap_branches = bool(args.overlap)
    if world > 1:  # identical replicas
        parallel.broadcast_parameters(model)
    ctc = nnet.CTCLoss(zero_infinity=True, assert_shorter=False)
    B = args.batch
    host = synth_inputs(args.model, B, None, seed=1234 + rank, pinned=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    params = [p for p in model.parameters()]
    opt = None
    if args.optimizer == "adam":
        opt = nnet.optimizers.Adam(model.parameters(), lr=nnet.schedulers.NoamDecayScheduler(10000, 360, 2), betas=(0.9, 0.98), eps=1e-9,
                                   weight_decay=1e-6, grad_max_norm=5.0)
        opt.flat()   # parameters become views of the flat buffer BEFORE anything is captured

    # N > 1: gradients are all-reduced bucket by bucket from inside the backward (hooks), on a communication stream
    buckets = None
    if world > 1:
        if args.sync_bn:
            ops.set_sync_batchnorm(True)
            if args.model == "AV":
                model.encoder.overlap_branches = False   # the statistics all-reduces of both branches share one communicator: single stream
        f = opt.flat() if opt is not None else None
        buckets = parallel.GradientBuckets(params, bucket_bytes=args.bucket_mb << 20, flat=f["g"] if f else None, offsets=f["offs"] if f else None)

    def fwd_bwd(d):
        """forward + 6 CTC losses + backward (+ the overlapped gradient all-reduce); leaves the gradients in p.grad"""
        for p in params:
            p.grad = None
        # every training step starts from weights the optimizer has just changed: the compute-dtype (bf16, kernel-layout) copies are
        # rebuilt inside the timed step even when --optimizer none leaves the fp32 masters untouched
        AF.invalidate_weights()
        outputs = model(model_inputs(args.model, d))
        if args.loss == "ctc":
            # fused CTC kernel with device-side lengths: no host sync, capturable in the CUDA graph
            loss = sum(ctc((d["labels"], d["llen"]), v) for v in outputs.values()) / len(outputs)
        else:
            loss = sum(v[0].float().mean() for v in outputs.values())
        loss.backward()
        if buckets is not None:
            buckets.finish()
        return loss.detach()

    static_grads = None   # gradient tensors the captured graph writes on every replay

    def allreduce_grads():
        pass   # done inside fwd_bwd (GradientBuckets)

    graph, static_loss = None, None

    def step(d):
        if graph is not None and d is resident:
            graph.replay()
            loss = static_loss
        else:
            loss = fwd_bwd(d)
        allreduce_grads()
        if opt is not None:
            opt.step(grads=static_grads if (graph is not None and d is resident) else None)
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (eager), optional CUDA-graph capture of forward+backward, warm-up again
    for _ in range(2):
        step(resident)
    use_graph = False
    if args.graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fwd_bwd(resident)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gobj = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gobj):
                static_loss = fwd_bwd(resident)
            graph, use_graph = gobj, True
            static_grads = [p.grad for p in params]
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed, running eager: {type(e).__name__}: {str(e)[:200]}", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    for _ in range(max(3, args.warmup)):
        step(resident)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: step(resident), args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host -> device copies + loss read back inside the timed region.  The H2D copy of the NEXT step's
    # inputs runs on a copy stream while the current step computes (double-buffered staging tensors, one H2D per step, the
    # step starts with a device-to-device copy into the inputs the captured graph reads).
    copy_stream = torch.cuda.Stream()
    staging = {k: torch.empty_like(resident[k]) for k in ("audio", "video", "labels")}

    def prefetch():
        copy_stream.wait_stream(torch.cuda.current_stream())   # the previous D2D has consumed the staging buffers
        with torch.cuda.stream(copy_stream):
            for k, v in staging.items():
                v.copy_(host[k], non_blocking=True)

    prefetch()

    def e2e_step():
        torch.cuda.current_stream().wait_stream(copy_stream)
        if graph is not None:
            for k, v in staging.items():
                resident[k].copy_(v)
            d = resident
        else:
            d = dict(resident)
            for k, v in staging.items():
                d[k] = v.clone()
        prefetch()
        return float(step(d).item())
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    h2d = sum(host[k].numel() * host[k].element_size() for k in staging)   # audio + video + labels (lengths are constant)

    # ---- roofline of the dominant kernel family (tcgen05 GEMM / implicit-GEMM conv): CUDA events around every launch of an eager
    # single-stream step (the two-stream schedule would fold the other branch's kernels into each window)
    if args.model == "AV":
        model.encoder.overlap_branches = False
    wg_overlap, AF.WGRAD_OVERLAP = AF.WGRAD_OVERLAP, False      # (weight gradients on the main stream too: one kernel at a time)
    prof = profile_gemm(lambda: fwd_bwd(resident), ops)
    AF.WGRAD_OVERLAP = wg_overlap
    if args.model == "AV":
        model.encoder.overlap_branches = bool(args.overlap)
    ops.reset_launch_count()
    fwd_bwd(resident)
    launches = ops.launch_count()   # kernels of this library in one eager step (a graph replay launches the same set)
    peaks = measured_peaks()
    parity = parity_check(args, model, resident, dev) if (rank == 0 and args.parity_check) else None

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * B / (ms_step / 1000.0)
        aug = (f"dropout {args.dropout:g} at every site + SpecAugment(2,27,5,0.05): the reference's training graph" if args.dropout > 0
               else "dropout 0, SpecAugment bypassed: parity graph")
        gemm_ms = prof["ms"]
        # algorithmic FLOPs: dense-contraction FLOPs of the reference graph per utterance x utterances per step;
        # every one of them runs in the tcgen05 kernels (strided dgrads counted at their algorithmic, not zero-inserted, size)
        algo_tf = GFLOP_PER_UTT[args.model] * B / 1000.0
        ach = algo_tf / (gemm_ms / 1000.0) if gemm_ms > 0 else 0.0
        whole = algo_tf / (ms_step / 1000.0)
        line = {
            "metric": "utterances/sec fwd+bwd (4s audio+video)", "value": value, "unit": "utterances/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{args.model} EffConfInterCTC fwd+bwd (train mode, {aug}, 6 CTC heads), per-GPU batch {B}, "
                                   f"64000 audio samples + {101 if args.model == 'AV' else 100}x88x88 video",
                       "global_batch": world * B, "parallelism": f"dp{world}", "bn": "SyncBatchNorm (all-reduced statistics)" if (world > 1 and args.sync_bn) else "local batch statistics", "optimizer": args.optimizer,
                       "allreduce": (f"{len(buckets.buckets)} fp32 bucket(s), NCCL all-reduce (AVG) launched from backward hooks on a communication stream, captured in the CUDA graph" if buckets is not None else None),
                       "streams": (3 if (args.model == "AV" and args.overlap and not (world > 1 and args.sync_bn)) else 1),   # caller + audio branch + video branch
                       "launch": "programmatic dependent launch (GEMM / attention / small kernels)" + (" on the caller's stream only" if args.model == "AV" and args.overlap else ""), "loss": args.loss, "cuda_graph": bool(use_graph),
                       "weights": "fp32 masters converted to bf16 kernel layout inside every timed step",
                       "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no flush needed",
                       "algorithmic_tflop_per_step": algo_tf, "achieved_tflops_whole_step": whole,
                       "frac_of_tensor_peak_whole_step": whole / peaks["sustained"], "frac_of_burst_peak_whole_step": whole / peaks["burst"]},
            "e2e": {"value": world * B / (ms_e2e / args.steps / 1000.0), "unit": "utterances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": ach / peaks["sustained"],
                         "frac_burst": ach / peaks["burst"], "peak_burst": peaks["burst"], "traffic": gemm_traffic(),
                         "kernel": "tcgen05 GEMM / implicit-GEMM conv kernels of csrc/gemm_tc.cu (gemm_tc_kernel, conv3x3_halo64, wgrad_halo64, wgrad_img, stem3d_*), "
                                   "all launches of one step; achieved = ALGORITHMIC dense-contraction FLOPs of the step (SURVEY 8d figure x batch) / their summed "
                                   "CUDA-event time in a single-stream eager step; traffic = their summed DRAM bytes (ncu)",
                         "launches_per_step": prof["n"], "ms_per_step_in_kernel": gemm_ms, "launched_tflop_per_step": prof["launched_tf"],
                         "families": prof["families"], "peak_source": peaks["which"] + " bf16_tflops_sustained (burst: bf16_tflops)"},
        }
        if parity is not None:
            line["parity"] = parity
        if args.shape_table:
            os.makedirs(os.path.dirname(os.path.abspath(args.shape_table)), exist_ok=True)
            json.dump({"workload": line["config"]["workload"], "shapes": prof["shapes"]}, open(args.shape_table, "w"), indent=1)
        if world == 1 and args.incumbent:
            try:
                line["incumbent"] = gpu_incumbent(args, dev, {k: v.detach().cpu() for k, v in model.state_dict().items()})
            except Exception as e:  # noqa: BLE001
                line["incumbent"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        if world == 1 and args.cpu_baseline:
            cb = cpu_baseline(args, steps=2, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # All ranks leave together and WITHOUT tearing NCCL down: rank 0 arrives here seconds after the others (it alone runs the
        # parity check), and destroying communicators whose kernels live inside captured CUDA graphs while peers are already gone
        # hung the 8-GPU run of this round until its time limit.  A hard exit after a final barrier cannot hang; a timer backs it up.
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(120.0, os._exit, args=(0,)).start()
        try:
            dist.barrier()
        finally:
            os._exit(0)


def profile_gemm(step_fn, ops):
    """CUDA-event pair around every tcgen05 GEMM / conv launch of an eager step.  The step is run three times behind a device-side
    sleep (so the host runs ahead of the GPU and an event window holds the kernel, not host launch latency) and every launch
    keeps the minimum of its three windows.  Returns summed time, launch count, launched and algorithmic FLOPs, a per-family
    split and the per-shape table."""
    from avec_b200 import _lib as L
    orig = ops._gemm
    stem_orig = (ops.stem3d_fwd, ops.stem3d_wgrad)
    runs = []

    def describe(a):
        launched = 2.0 * a.M * a.N * a.K
        if a.mode == L.GEMM_PLAIN:
            fam = "linear wgrad" if a.epi == L.EPI_ACCUM else ("linear dgrad" if a.sbk != 1 else "linear fwd")
            return fam, launched, launched, f"M{a.M} N{a.N} K{a.K}"
        g = a.g
        taps = g.KT * g.KH * g.KW
        sites_out = g.N * g.To * g.Ho * g.Wo
        algo = 2.0 * sites_out * taps * g.C * g.Co       # strided dgrads: the algorithmic count, not the zero-inserted one
        fam = {L.GEMM_CONV_FWD: "conv fwd", L.GEMM_CONV_DGRAD: "conv dgrad", L.GEMM_CONV_WGRAD: "conv wgrad"}[a.mode]
        return fam, launched, algo, f"N{g.N} {g.Hi}x{g.Wi} C{g.C}->{g.Co} k{g.KH} s{g.sh}"

    for _ in range(3):
        recs = []

        def hooked(a, may_decline=False, recs=recs):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launched = orig(a, may_decline)
            e1.record()
            if launched is not False:      # (a GEMM with a fused dropout may decline: the un-fused pair follows and is recorded)
                recs.append((e0, e1) + describe(a))
            return launched

        def timed_stem(fn, name, recs=recs):
            def wrapper(x, *a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(x, *a, **k)
                e1.record()
                B_, T_, H_, W_ = x.shape[0], x.shape[1], x.shape[2], x.shape[3]
                f = 2.0 * B_ * T_ * (H_ // 2) * (W_ // 2) * 64 * 245
                recs.append((e0, e1, name, f, f, f"N{B_ * T_} {H_}x{W_} C1->64 k(5,7,7)"))
                return out
            return wrapper
        ops._gemm = hooked
        ops.stem3d_fwd, ops.stem3d_wgrad = timed_stem(stem_orig[0], "stem3d fwd"), timed_stem(stem_orig[1], "stem3d wgrad")
        try:
            torch.cuda._sleep(int(0.4 * 1.9e9))      # ~0.4 s head start: the eager step's host work (~0.15 s) never catches up with the GPU
            step_fn()
            torch.cuda.synchronize()
        finally:
            ops._gemm = orig
            ops.stem3d_fwd, ops.stem3d_wgrad = stem_orig
        runs.append([(e0.elapsed_time(e1),) + tuple(rest) for e0, e1, *rest in recs])
    n = min(len(r) for r in runs)
    best = [min(r[i][0] for r in runs) for i in range(n)]
    fam, shapes = {}, {}
    for i in range(n):
        _, f, launched, algo, shape = runs[0][i]
        a = fam.setdefault(f, {"ms": 0.0, "launches": 0, "algorithmic_tflop": 0.0})
        a["ms"] += best[i]; a["launches"] += 1; a["algorithmic_tflop"] += algo / 1e12
        b = shapes.setdefault((f, shape), {"family": f, "shape": shape, "ms": 0.0, "launches": 0, "algorithmic_tflop": 0.0})
        b["ms"] += best[i]; b["launches"] += 1; b["algorithmic_tflop"] += algo / 1e12
    for a in list(fam.values()) + list(shapes.values()):
        a["tflops"] = a["algorithmic_tflop"] / (a["ms"] / 1000.0) if a["ms"] > 0 else 0.0
    if os.environ.get("AVEC_BENCH_VERBOSE"):
        for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:16]:
            print(f"[bench] {k}: {v['ms']:.3f} ms x{v['launches']} {v['tflops']:.0f} TFLOP/s", file=sys.stderr)
    return {"ms": sum(best), "n": n, "launched_tf": sum(runs[0][i][2] for i in range(n)) / 1e12,
            "families": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in fam.items()},
            "shapes": sorted(shapes.values(), key=lambda v: -v["ms"])}


if __name__ == "__main__":
    main()
