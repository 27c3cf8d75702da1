"""Zoo models with the reference's constructor arguments and forward(inputs) -> dict convention
(reference nnet/models_zoo.py:64-182).  Only the encoder hot path is re-implemented; `Model` here is a light nn.Module
base (compile / forward / losses) - the reference's training runtime (nnet/model.py fit/evaluate/save/...) is out of
scope and keeps working by patching these encoders into it (avec_b200.patch_reference())."""
import torch
import torch.nn as nn

from .. import functional as AF
from . import networks
from .losses import CTCLoss


class Model(nn.Module):
    def __init__(self, name="model"):
        super().__init__()
        self.name = name
        self.compiled = False
        self.losses = None
        self.loss_weights = None

    # the zoo's default loss weights, mapped to the outputs by position (lists) or key (dicts): models_zoo.py:81,131,166
    default_loss_weights = None

    def compile(self, losses=None, loss_weights=None, optimizer=None, metrics=None, decoders=None, grad_max_norm=None, ema_tau=None):
        """same arguments as the zoo models' compile (models_zoo.py:78-99,128-149,163-184): optimizer="Adam" builds the reference's
        default (Noam schedule 10000 / 360 / 2, betas (0.9, 0.98), eps 1e-9, L2 weight decay 1e-6) as the fused flat-buffer
        optimizer; grad_max_norm / ema_tau (Model arguments in the reference, model.py:378-404) are folded into the same launch."""
        from . import optimizers, schedulers
        self.losses = losses if losses is not None else CTCLoss()
        self.loss_weights = loss_weights if loss_weights is not None else self.default_loss_weights
        if optimizer == "Adam":
            lr = schedulers.NoamDecayScheduler(warmup_steps=10000, dim_decay=360, val_factor=2)
            optimizer = optimizers.Adam(params=self.parameters(), lr=lr, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6,
                                        grad_max_norm=grad_max_norm, ema_tau=ema_tau)
        self.optimizer, self.metrics, self.decoders = optimizer, metrics, decoders
        self.compiled = True

    def train_step(self, inputs, targets):
        """forward + weighted CTC losses + backward + optimizer step (the arithmetic of Model.train_step, model.py:342-407,
        without grad scaler / accumulation / logging).  Returns the detached total loss (no host sync)."""
        assert self.compiled and self.optimizer is not None, "compile(optimizer=...) first"
        outputs = self(inputs)
        loss = self.compute_loss(outputs, targets)
        for p in self.parameters():
            p.grad = None
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def _scope(self, device):
        """opens the forward pass (zero arena, dropout sites, RNG step): avec_b200.functional.forward_scope"""
        return AF.forward_scope(self, device)

    def compute_loss(self, outputs, targets):
        """sum_k w_k * CTC(outputs[k]) with weights mapped to the outputs by position for lists (model.py:217) or by
        key for dicts - the reference's forward_model loss bookkeeping (model.py:275-287)."""
        keys = list(outputs.keys())
        w = self.loss_weights
        if w is None:
            w = [1.0] * len(keys)
        if isinstance(w, dict):
            w = [w[k] for k in keys]
        loss_fn = self.losses if self.losses is not None else CTCLoss()
        total = 0.0
        for k, wk in zip(keys, w):
            total = total + wk * loss_fn(targets, outputs[k])
        return total


class AudioEfficientConformerInterCTC(Model):
    default_loss_weights = [0.5 / 4, 0.5 / 4, 0.5 / 4, 0.5 / 4, 0.5]

    def __init__(self, vocab_size=256, att_type="patch", interctc_blocks=[3, 6, 10, 13]):
        super().__init__(name="Audio Efficient Conformer Inter CTC")
        self.encoder = networks.AudioEfficientConformerEncoder(vocab_size=vocab_size, att_type=att_type, interctc_blocks=interctc_blocks)

    def forward(self, inputs):
        x, lengths = inputs
        with self._scope(x.device):
            x, lengths, interctc_outputs = self.encoder(x, lengths)
        outputs = {"outputs": [x, lengths]}
        outputs.update(interctc_outputs)
        return outputs


class VisualEfficientConformerInterCTC(Model):
    default_loss_weights = [0.5 / 3, 0.5 / 3, 0.5 / 3, 0.5]

    def __init__(self, vocab_size=256, interctc_blocks=[3, 6, 9], test_augments=None):
        super().__init__(name="Visual Efficient Conformer Inter CTC")
        self.encoder = networks.VisualEfficientConformerEncoder(vocab_size=vocab_size, interctc_blocks=interctc_blocks)
        self.test_augments = test_augments if isinstance(test_augments, list) else [test_augments] if test_augments is not None else None

    def forward(self, inputs):
        """Evaluation with test-time augmentation (models_zoo.py:113-122; the VO config passes RandomHorizontalFlip(p=1)): the
        reference runs the encoder once per augment; here the clean clip and its augmented copies go through ONE encoder pass as a
        batch of (1 + A) * B clips (BatchNorm uses running statistics in eval(), so batching is exact) and the logits are
        re-stacked to the reference's (B, 1 + A, T, V) / (B, 1 + A)."""
        video, video_lengths = inputs
        assert not (self.training and self.test_augments is not None), "Training requires setting test_time_aug to False / test_augments to None"
        with self._scope(video.device):
            if not self.training and self.test_augments is not None:
                v = video.permute(0, 4, 1, 2, 3) if video.shape[-1] == 1 else video          # (B, 1, T, H, W) as the reference passes it
                clips = [v] + [aug(v) for aug in self.test_augments]
                n, B = len(clips), v.shape[0]
                x, lengths, interctc_outputs = self.encoder(torch.cat(clips, dim=0).contiguous(), video_lengths.repeat(n))
                x = x.view(n, B, *x.shape[1:]).transpose(0, 1).contiguous()
                lengths = lengths.view(n, B).t().contiguous()
                interctc_outputs = {k: [val[0][:B], val[1][:B]] for k, val in interctc_outputs.items()}   # the clean clip's heads
            else:
                x, lengths, interctc_outputs = self.encoder(video, video_lengths)
        outputs = {"outputs": [x, lengths]}
        outputs.update(interctc_outputs)
        return outputs


class AudioVisualEfficientConformerInterCTC(Model):
    default_loss_weights = {"v_ctc_2": 0.5 / 3, "v_ctc_5": 0.5 / 3, "a_ctc_7": 0.5 / 3, "a_ctc_10": 0.5 / 3, "f_ctc_1": 0.5 / 3, "outputs": 0.5}

    def __init__(self, vocab_size=256, v_interctc_blocks=[3, 6], a_interctc_blocks=[8, 11], f_interctc_blocks=[2]):
        super().__init__(name="Audio-Visual Efficient Conformer Inter CTC")
        self.encoder = networks.AudioVisualEfficientConformerEncoder(
            vocab_size=vocab_size, v_interctc_blocks=v_interctc_blocks, a_interctc_blocks=a_interctc_blocks,
            f_interctc_blocks=f_interctc_blocks)

    def forward(self, inputs):
        video, video_len, audio, audio_len = inputs
        with self._scope(video.device):
            x, lengths, interctc_outputs = self.encoder(video, video_len, audio, audio_len)
        outputs = {"outputs": [x, lengths]}
        outputs.update(interctc_outputs)
        return outputs
