"""CPU suite: the training-step oracle (oracle/train_oracle.py, oracle/restate.py with dropout multipliers) pinned against
(a) the published Philox4x32-10 known-answer vectors (Random123 kat_vectors) and (b) fixtures produced by the UNMODIFIED
reference with its random draws replaced by that Philox stream (oracle/make_golden_train.py)."""
import numpy as np
import pytest
import torch

from common import block_G, check_close, check_grad_fingerprint
from conftest import load_golden
import seeded
from oracle import restate, train_oracle as TO

KAT = [  # counter (4), key (2), output (4): Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    got = tuple(int(w) for w in TO.philox4x32_10(*ctr, *key))
    assert got == want


def test_dropout_keep_statistics_and_determinism():
    k1 = TO.dropout_keep(7, 3, 11, 4096, 180, 0.1)
    k2 = TO.dropout_keep(7, 3, 11, 4096, 180, 0.1)
    assert (k1 == k2).all() and k1.shape == (4096, 180)
    assert abs(1.0 - k1.mean() - 0.1) < 2e-3
    assert (TO.dropout_keep(7, 4, 11, 4096, 180, 0.1) != k1).mean() > 0.1       # another step: another mask
    assert (TO.dropout_keep(7, 3, 12, 4096, 180, 0.1) != k1).mean() > 0.1       # another site: another mask
    assert TO.dropout_keep(7, 3, 11, 64, 180, 0.0).all()


def test_spec_augment_restatement_matches_reference():
    fix = load_golden("train_spec_augment.pt")
    B, M, F = fix["shape"]
    mel = seeded.randn("specaug.mel", (B, M, F), fix["mel_seed"])
    got = TO.spec_augment(mel.transpose(1, 2).contiguous().numpy(), fix["lengths"].tolist(), fix["seed"], fix["step"], fix["site"],
                          *fix["params"])
    want = fix["out"].transpose(1, 2).numpy()
    assert np.array_equal(got, want)
    assert 0.05 < fix["zero_frac"] < 0.9


@pytest.mark.parametrize("tag", ["s1_patch_T20", "s2_down_T12", "s1_grouped3_T20"])
def test_block_restatement_with_dropout_matches_reference(tag):
    fix = load_golden(f"train_block_dropout_{tag}.pt")
    cfg, p = fix["cfg"], fix["p"]
    from common import make_block
    blk, sd = make_block(cfg)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    x = seeded.randn(tag + ".x", (cfg["B"], cfg["T"], cfg["D"]), cfg["seed"]).requires_grad_(True)
    drop = [torch.from_numpy(TO.dropout_scale_mask(fix["rng_seed"], fix["step"], i + 1, shp[0] * shp[1], shp[2], p)).view(shp)
            for i, shp in enumerate(fix["site_shapes"])]
    P = 3 if cfg["att"] == "patch" else 1
    y, _, _ = restate.conformer_block(x, sd, fix["lengths"], 4, P, cfg["stride"], True, G=block_G(cfg), drop=drop)
    gy = seeded.randn(tag + ".gy", tuple(y.shape), cfg["seed"])
    (y * gy).sum().backward()
    check_close("y", y, fix["y"], 1e-4, 1e-5)
    check_close("dx", x.grad, fix["dx"], 1e-4, 1e-5)
    for k, fp in fix["grads"].items():
        check_grad_fingerprint(k, sd[k].grad, fp, 1e-3, 1e-4)


def test_adam_restatement_matches_reference():
    fix = load_golden("train_adam.pt")
    h = fix["hyper"]
    p, m, v, ema = fix["p0"].numpy(), np.zeros(4096), np.zeros(4096), fix["p0"].numpy()
    for i, g in enumerate(fix["grads"]):
        lr = TO.noam_lr(i + 1, h["warmup_steps"], h["dim_decay"], h["val_factor"])
        assert abs(lr - fix["lr"][i]) <= 1e-6 * lr
        p, m, v, ema = TO.adam_step(p, g.numpy(), m, v, i + 1, lr, *h["betas"], h["eps"], h["weight_decay"], h["max_norm"], ema,
                                    h["ema_tau"])
        np.testing.assert_allclose(p, fix["p"][i].numpy(), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(ema, fix["ema"][i].numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(m, fix["m"].numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(v, fix["v"].numpy(), rtol=2e-5, atol=1e-9)


def test_greedy_restatement_matches_reference():
    fix = load_golden("train_greedy.pt")
    assert TO.greedy_decode(fix["logits"].numpy(), fix["lengths"].tolist(), 0) == fix["tokens"]
    assert fix["tokens"][1] == [] and fix["tokens"][4] == []


def test_schedulers_match_reference_formulas():
    """avec_b200.nnet.schedulers (host side of the on-device learning-rate schedule) against the reference's classes when the
    reference tree is importable (authoring container), and against the closed form otherwise"""
    from avec_b200.nnet import schedulers as S
    mine = S.NoamDecayScheduler(warmup_steps=10000, dim_decay=360, val_factor=2)
    steps = [1, 2, 17, 9999, 10000, 10001, 250000]
    want = [2 * 360 ** -0.5 * min(s * 10000 ** -1.5, s ** -0.5) for s in steps]
    from oracle import ref_import
    if ref_import.available():
        ref = ref_import.import_reference()
        r = ref.schedulers.NoamDecayScheduler(warmup_steps=10000, dim_decay=360, val_factor=2)
        want = [float(r.get_val_step(s)) for s in steps]
        assert float(ref.schedulers.ConstantScheduler(val=3e-4).get_val_step(5)) == S.ConstantScheduler(3e-4).get_val_step(5)
    for s, w in zip(steps, want):
        assert abs(mine.get_val_step(s) - w) <= 1e-12 + 1e-9 * w
        assert abs(TO.noam_lr(s) - w) <= 1e-12 + 1e-9 * w
    a, b = mine.device_params()                      # lr = a * min(t * b^-1.5, t^-0.5) evaluated by adam_kernel
    assert abs(a * min(17 * b ** -1.5, 17 ** -0.5) - mine.get_val_step(17)) < 1e-15
    assert mine.step() == mine.get_val_step(1) and int(mine.model_step) == 1


def test_video_augment_restatement_matches_reference():
    """oracle/train_oracle.video_augment == torchvision RandomCrop + RandomHorizontalFlip + the reference's TimeMaskSecond fed with the
    same draws (fixture generated by those modules themselves, oracle/make_golden_train.py video_augment_case)"""
    fix = load_golden("train_video_augment.pt")
    video = seeded.randn("videoaug.x", fix["shape"], fix["video_seed"]).clamp(-1, 1).numpy()
    out = TO.video_augment(video, fix["lengths"].tolist(), fix["seed"], fix["step"], fix["site"])
    assert np.abs(out[:, ::3, ::4, ::4] - fix["out_sub"].numpy()).max() <= 1e-6
    assert abs(float(out.astype(np.float64).sum()) - fix["out_sum"]) <= 1e-3
    B, T, Hi, Wi = fix["shape"]
    for b in range(B):
        got = TO.video_draws(fix["seed"], fix["step"], fix["site"], b, int(fix["lengths"][b]), Hi, Wi, 88, 88)
        assert got == fix["draws"][b]
