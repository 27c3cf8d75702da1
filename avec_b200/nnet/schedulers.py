"""Learning-rate schedules with the reference's interface (nnet/schedulers.py:24-137).  The fused optimizer evaluates
the schedule ON THE DEVICE from its step counter (csrc/train.cu adam_kernel); get_val_step is the host-side formula."""
import torch
import torch.nn as nn


class Scheduler(nn.Module):
    lr_mode = None                       # code understood by avec_adam_step (0 constant, 1 Noam)

    def __init__(self):
        super().__init__()
        self.model_step = torch.tensor(0)

    def step(self):
        self.model_step += 1
        return self.get_val()

    def get_val(self):
        return self.get_val_step(self.model_step)

    def get_val_step(self, step):
        return None

    def device_params(self):
        raise NotImplementedError


class ConstantScheduler(Scheduler):
    lr_mode = 0

    def __init__(self, val):
        super().__init__()
        self.val = val

    def get_val_step(self, step):
        return self.val

    def device_params(self):
        return float(self.val), 1.0


class NoamDecayScheduler(Scheduler):
    """val_factor * dim_decay^-0.5 * min(step * warmup_steps^-1.5, step^-0.5)  (schedulers.py:120-137)"""
    lr_mode = 1

    def __init__(self, warmup_steps, dim_decay, val_factor):
        super().__init__()
        self.warmup_steps, self.dim_decay, self.val_factor = warmup_steps, dim_decay, val_factor

    def get_val_step(self, step):
        step = float(step)
        return self.val_factor * self.dim_decay ** -0.5 * min(step * self.warmup_steps ** -1.5, step ** -0.5)

    def device_params(self):
        return float(self.val_factor * self.dim_decay ** -0.5), float(self.warmup_steps)


def adopt(lr):
    """Scheduler of this package for `lr`: a float, one of our schedulers, or the reference's own scheduler object (duck-typed by
    class name: after avec_b200.patch_reference() the zoo models hand the reference's NoamDecayScheduler to the fused Adam,
    models_zoo.py:172-174).  The host-side `model_step` tensor is shared with the adopted object."""
    if isinstance(lr, Scheduler):
        return lr
    if hasattr(lr, "model_step") and hasattr(lr, "get_val_step"):
        name = lr.__class__.__name__
        if name == "NoamDecayScheduler":
            s = NoamDecayScheduler(lr.warmup_steps, lr.dim_decay, lr.val_factor)
        elif name == "ConstantScheduler":
            s = ConstantScheduler(lr.val)
        else:
            raise NotImplementedError(f"avec_b200 fused Adam: learning-rate schedule {name} is not implemented on the device")
        s.model_step = lr.model_step
        return s
    return ConstantScheduler(val=lr)
