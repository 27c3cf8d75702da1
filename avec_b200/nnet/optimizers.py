"""Fused optimizer step with the reference's interface (nnet/optimizers.py:61-93: torch Adam whose `lr` is a Scheduler
stepped inside `step()`; nnet/model.py:378-407: global-norm clipping before the step, EMA after it).

Parameters, gradients, both moments (and the EMA copy) live in FLAT fp32 buffers; every nn.Parameter becomes a view of
the flat parameter buffer, so the model, checkpoints and the flat NCCL gradient bucket (avec_b200.parallel) see the same
memory.  One step = gather of the gradients (multi-tensor copy, skipped for gradients that already are views of the
flat buffer) + avec_sumsq (only when clipping) + avec_counter_advance + avec_adam_step: 3-4 launches instead of the
~5500 of torch.optim.Adam + clip_grad_norm_ + the per-tensor EMA loop on the 1102 tensors of the AV model.  Step count
and learning rate stay on the device, so the whole step is capturable in a CUDA graph."""
import torch

from .. import ops
from . import schedulers


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-08, weight_decay=0, amsgrad=False, grad_max_norm=None,
                 ema_tau=None):
        assert not amsgrad, "amsgrad is not used by the reference configs"
        self.scheduler = schedulers.adopt(lr)
        super().__init__(params, dict(lr=0.0, betas=betas, eps=eps, weight_decay=weight_decay))
        assert len(self.param_groups) == 1, "one parameter group (the reference passes model.parameters())"
        self.grad_max_norm, self.ema_tau = grad_max_norm, ema_tau
        self._flat = None

    # ---- flat buffers -------------------------------------------------------------------------------------------
    def _build(self):
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("avec_b200.nnet.optimizers.Adam needs CUDA parameters (no CPU fallback)")
        offs, n = [], 0
        for p in ps:
            assert p.dtype == torch.float32 and p.device == dev
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4                    # 16-byte aligned views
        f = {"params": ps, "offs": offs, "n": n}
        f["p"] = torch.zeros(n, device=dev)
        for k in ("g", "m", "v"):
            f[k] = torch.zeros(n, device=dev)
        f["ema"] = torch.zeros(n, device=dev) if self.ema_tau is not None else None
        f["step"] = torch.zeros(1, device=dev, dtype=torch.int64)
        f["sumsq"] = torch.zeros(1, device=dev)
        f["info"] = torch.zeros(2, device=dev)
        f["gviews"] = []
        with torch.no_grad():
            for p, o in zip(ps, offs):
                view = f["p"][o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                f["gviews"].append(f["g"][o:o + p.numel()].view_as(p))
                self.state[p] = {"step": f["step"], "exp_avg": f["m"][o:o + p.numel()].view_as(p),
                                 "exp_avg_sq": f["v"][o:o + p.numel()].view_as(p)}
            if f["ema"] is not None:
                f["ema"].copy_(f["p"])
        self._flat = f
        return f

    def flat(self):
        return self._flat if self._flat is not None else self._build()

    def grad_views(self):
        """views of the flat gradient buffer, one per parameter: point `p.grad` (or a CUDA graph's static gradients, or the NCCL
        bucket) at them and step() needs no gather"""
        return self.flat()["gviews"]

    def ema_parameters(self):
        f = self.flat()
        return [f["ema"][o:o + p.numel()].view_as(p) for p, o in zip(f["params"], f["offs"])] if f["ema"] is not None else None

    # ---- step ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, grads=None):
        """grads: optional list (one per parameter) used instead of p.grad, e.g. the static outputs of a captured backward.
        Parameters without a gradient are skipped entirely (no weight decay, no moment update), as torch.optim.Adam does:
        the fused kernel is then launched once per contiguous run of parameters that do have one."""
        assert closure is None
        from .. import functional as AF
        f = self.flat()
        if grads is not None and len(grads) != len(f["params"]):
            raise ValueError(f"grads has {len(grads)} entries for {len(f['params'])} trainable parameters")
        src, dst, runs = [], [], []
        for i, (p, gv, o) in enumerate(zip(f["params"], f["gviews"], f["offs"])):
            g = grads[i] if grads is not None else p.grad
            if g is None:
                gv.zero_()                      # keeps the global norm exact; the range itself is not stepped
                continue
            end = o + (p.numel() + 3) // 4 * 4
            if runs and runs[-1][1] == o:
                runs[-1][1] = end
            else:
                runs.append([o, end])
            if g.data_ptr() != gv.data_ptr():
                src.append(g)
                dst.append(gv)
        if src:
            torch._foreach_copy_(dst, src)
        grp = self.param_groups[0]
        sumsq = None
        if self.grad_max_norm is not None:
            f["sumsq"].zero_()
            sumsq = ops.sumsq(f["g"], f["sumsq"])
        L = ops.L
        L.check(L.load().avec_counter_advance(f["step"].data_ptr(), ops._stream()), "avec_counter_advance")
        self.scheduler.model_step += 1
        lr_a, lr_b = self.scheduler.device_params()
        for lo, hi in runs:
            ops.adam_step(f["p"][lo:hi], f["g"][lo:hi], f["m"][lo:hi], f["v"][lo:hi], f["step"], self.scheduler.lr_mode, lr_a, lr_b,
                          grp["betas"], grp["eps"], grp["weight_decay"], sumsq, float(self.grad_max_norm or 0.0),
                          f["ema"][lo:hi] if f["ema"] is not None else None, float(self.ema_tau or 0.0), f["info"])
        grp["lr"] = self.scheduler.get_val()
        AF.invalidate_weights()     # the parameters changed behind torch's version counters: drop the cached bf16 copies
        return None

    def last_info(self):
        """{"lr", "grad_norm"} of the last step, read back from the device (one sync; for logging only)"""
        lr, gn = self.flat()["info"].tolist()
        return {"lr": lr, "grad_norm": gn}

    def sync_host_step(self):
        """the step counter lives on the device (a captured CUDA graph advances it without running this Python code): copy it
        back into the host-side scheduler / param_groups (one sync; call before logging or checkpointing)"""
        step = int(self.flat()["step"].item())
        self.scheduler.model_step.fill_(step)
        self.param_groups[0]["lr"] = self.scheduler.get_val() if step > 0 else 0.0
        return step

    # ---- checkpoints: torch.optim.Adam layout + "model_step" (optimizers.py:75-91) ---------------------------------
    def state_dict(self):
        """torch.optim.Adam's layout: every parameter gets its OWN float32 scalar `step` (torch increments it in place per
        parameter; a shared tensor would be bumped once per parameter after a reload into the reference's Adam) and cloned
        moments; `model_step` is read from the device counter, so it is right even when the step ran inside a CUDA graph."""
        step = self.sync_host_step()
        sd = super().state_dict()
        sd["state"] = {k: {"step": torch.tensor(float(step), dtype=torch.float32), "exp_avg": st["exp_avg"].detach().clone(),
                           "exp_avg_sq": st["exp_avg_sq"].detach().clone()} for k, st in sd["state"].items()}
        sd["model_step"] = self.scheduler.model_step.clone()
        return sd

    def load_state_dict(self, state_dict):
        from .. import functional as AF
        state_dict = dict(state_dict)
        step = int(state_dict.pop("model_step"))
        self.scheduler.model_step.fill_(step)
        f = self.flat()
        views = {id(p): (self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) for p in f["params"]}
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for p in f["params"]:
                st = self.state.get(p, {})
                m, v = views[id(p)]
                if "exp_avg" in st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                self.state[p] = {"step": f["step"], "exp_avg": m, "exp_avg_sq": v}
            f["step"].fill_(step)
        AF.invalidate_weights()
