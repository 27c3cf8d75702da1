"""The drop-in boundary EXECUTED: the reference's own `nnet.Model` runtime and `main.py`
launcher running on the avec_b200 encoders after `avec_b200.patch_reference()`.

Needs the unmodified reference tree (/root/reference in the authoring container, baseline/_ref on the GPU box - put there by
oracle/install_ref.sh, git-ignored).  Without it these tests skip: nothing else in the GPU suite reads the reference."""
import glob
import os
import subprocess
import sys

import pytest
import torch

import avec_b200
import seeded
from avec_b200 import ops
from common import rel_err
from oracle import ref_import

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (oracle/install_ref.sh)")]
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(B=4, Ls=5120, seed=5):
    Tv = Ls // 640 + 1
    audio = seeded.randn("dropin.audio", (B, Ls), seed, 0.1)
    video = seeded.randn("dropin.video", (B, Tv, 88, 88, 1), seed).clamp(-1, 1)
    alen = torch.tensor([Ls - 640 * (i % 3) for i in range(B)])
    vlen = alen // 640 + 1
    for b in range(B):
        audio[b, int(alen[b]):] = 0.0
        video[b, int(vlen[b]):] = 0.0
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(1, 256, (B, 6), generator=g)
    return [video, vlen, audio, alen], (labels, torch.full((B,), 6))


def _to(dev, xs):
    return type(xs)(x.to(dev) for x in xs)     # targets stay a TUPLE: a list would be mapped output by output (model.py:212-216)


class _NoAug(torch.nn.Module):
    def forward(self, x, lengths=None):
        return x


def _zoo_model(ref, seed=3):
    m = ref.AudioVisualEfficientConformerInterCTC(vocab_size=256)
    m.compile(losses=ref.CTCLoss(zero_infinity=True, assert_shorter=False))
    return m


def test_reference_model_runtime_on_patched_encoders(tmp_path):
    """reference zoo Model + reference train_step / save / load, avec_b200 encoders underneath:
    fp32 losses equal the unpatched reference's on the same weights; two bf16 training steps run through Model.train_step;
    a checkpoint written by Model.save loads back into an UNPATCHED reference model (same keys, torch-Adam state layout)."""
    ref = ref_import.import_reference()
    inputs, targets = _batch()
    # ---- the unpatched reference on the GPU (eager fp32, dropout off): the values to match
    torch.manual_seed(0)
    plain = _zoo_model(ref)
    sd = {k: v.clone() for k, v in plain.state_dict().items()}
    ref_import.zero_dropout(plain)
    plain.encoder.audio_encoder.spec_augment = _NoAug()
    plain = plain.to(DEV).train()
    want, _, _, _ = plain.forward_model(_to(DEV, inputs), _to(DEV, targets), compute_metrics=False)
    want = {k: float(v) for k, v in want.items()}
    try:
        avec_b200.patch_reference(ref)
        assert ref.networks.AudioVisualEfficientConformerEncoder is avec_b200.nnet.AudioVisualEfficientConformerEncoder
        fast = _zoo_model(ref)
        assert isinstance(fast, ref.model.Model) and isinstance(fast.optimizer, avec_b200.nnet.optimizers.Adam)
        assert list(fast.state_dict().keys()) == list(sd.keys())
        fast.load_state_dict(sd)
        fast = fast.to(DEV).train()
        avec_b200.nnet.zero_dropout(fast)
        # fp32 (no autocast -> "auto" picks the fp32 parity kernels): the reference's loss bookkeeping on our logits
        got, _, _, _ = fast.forward_model(_to(DEV, inputs), _to(DEV, targets), compute_metrics=False)
        for k, v in want.items():
            assert abs(float(got[k]) - v) <= 1e-3 * abs(v) + 1e-4, f"{k}: {float(got[k])} vs reference {v}"
        # two training steps through the reference's own train_step in bf16 autocast (dropout 0.1 + SpecAugment back on)
        for mod in fast.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.1
            if isinstance(mod, avec_b200.nnet.SpecAugment):
                mod.enabled = True
        scaler = torch.cuda.amp.GradScaler(enabled=False)
        w0 = fast.encoder.head.weight.detach().clone()
        site_counts = []
        for _ in range(2):
            losses, _, acc = fast.train_step(_to(DEV, inputs), _to(DEV, targets), torch.bfloat16, scaler, 1, 0, False)
            assert acc == 0 and torch.isfinite(losses["loss"])
            site_counts.append(ops.RNG.site)
        assert int(fast.model_step) == 2 and site_counts[0] == site_counts[1] > 100     # sites restart every forward
        assert int(ops.RNG.get(DEV)[1]) >= 2                                            # the RNG step advanced per forward
        assert not torch.equal(w0, fast.encoder.head.weight)
        assert all(bool(torch.isfinite(q).all()) for q in fast.parameters()), "non-finite parameter after the fused optimizer steps"
        # checkpoint round trip through the reference's save / load
        path = str(tmp_path / "checkpoints_epoch_1_step_2.ckpt")
        fast.save(path, save_optimizer=True)
    finally:
        avec_b200.unpatch_reference(ref)
        avec_b200.set_compute_dtype(torch.bfloat16)
    again = _zoo_model(ref)           # unpatched reference model + the reference's own torch Adam
    assert not isinstance(again.optimizer, avec_b200.nnet.optimizers.Adam)
    again = again.to(DEV)
    again.load(path)
    assert int(again.model_step) == 2
    assert rel_err(again.encoder.head.weight, fast.encoder.head.weight) == 0.0
    steps = {float(st["step"]) for st in again.optimizer.state_dict()["state"].values()}
    assert steps == {2.0}
    # one step of the REFERENCE optimizer from the reloaded state: every per-parameter step counter advances by exactly one
    for p in again.parameters():
        p.grad = torch.zeros_like(p)
    again.optimizer.step()
    assert {float(st["step"]) for st in again.optimizer.state_dict()["state"].values()} == {3.0}


def _run_main(extra, tmp_path, timeout=900):
    root = ref_import.reference_root()
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "oracle", "ref_stubs"), ROOT, env.get("PYTHONPATH", "")])
    env["AVEC_SYNTH_CALLBACKS"] = str(tmp_path)
    cmd = [sys.executable, os.path.join(root, "main.py"), "-c", "configs/synth/AV.py", "--steps_per_epoch", "2", "--step_log_period", "1"] + extra
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_reference_main_py_runs_unchanged(tmp_path):
    """`python main.py -c configs/synth/AV.py`: the reference's launcher, fit loop, evaluation and checkpointing, unmodified"""
    r = _run_main([], tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert glob.glob(str(tmp_path / "AV" / "checkpoints_epoch_1_step_2.ckpt")), r.stdout[-1500:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_reference_main_py_distributed(tmp_path):
    """`main.py -d`: mp.spawn, NCCL process group, Model.distribute_strategy (DDP over SyncBatchNorm-converted modules,
    nnet/model.py:59-61) around the two-stream AV encoder"""
    r = _run_main(["-d", "--world_size", "2", "--dist_addr", "127.0.0.1"], tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert glob.glob(str(tmp_path / "AV" / "checkpoints_epoch_1_step_2.ckpt")), r.stdout[-1500:]
