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
        self.scheduler = lr if isinstance(lr, schedulers.Scheduler) else schedulers.ConstantScheduler(val=lr)
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
        """grads: optional list (one per parameter) used instead of p.grad, e.g. the static outputs of a captured backward"""
        assert closure is None
        f = self.flat()
        if grads is not None and len(grads) != len(f["params"]):
            raise ValueError(f"grads has {len(grads)} entries for {len(f['params'])} trainable parameters")
        src, dst = [], []
        for i, (p, gv) in enumerate(zip(f["params"], f["gviews"])):
            g = grads[i] if grads is not None else p.grad
            if g is None:
                gv.zero_()
            elif g.data_ptr() != gv.data_ptr():
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
        ops.adam_step(f["p"], f["g"], f["m"], f["v"], f["step"], self.scheduler.lr_mode, lr_a, lr_b, grp["betas"], grp["eps"],
                      grp["weight_decay"], sumsq, float(self.grad_max_norm or 0.0), f["ema"],
                      float(self.ema_tau or 0.0), f["info"])
        grp["lr"] = self.scheduler.get_val()
        return None

    def last_info(self):
        """{"lr", "grad_norm"} of the last step, read back from the device (one sync; for logging only)"""
        lr, gn = self.flat()["info"].tolist()
        return {"lr": lr, "grad_norm": gn}

    # ---- checkpoints: torch.optim.Adam layout + "model_step" (optimizers.py:75-91) ---------------------------------
    def state_dict(self):
        self.flat()
        sd = super().state_dict()
        sd["model_step"] = self.scheduler.model_step
        return sd

    def load_state_dict(self, state_dict):
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
