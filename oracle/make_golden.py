"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference (burchim/AVEC, imported from
/root/reference through oracle/ref_import.py) on CPU in fp32 with seeded parameters and inputs (tests/seeded.py),
dropout forced to 0 and SpecAugment bypassed.  The GPU box has no reference tree:
these fixtures are what the `-m gpu` parity tests compare the CUDA path against.

    python oracle/make_golden.py            # (re)writes every fixture; ~1 minute on 8 cores
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_import  # noqa: E402
import seeded  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ref = ref_import.import_reference()


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def grads_of(module, max_n=4096):
    return {k: seeded.subsample(p.grad, max_n) for k, p in module.named_parameters() if p.grad is not None}


def mask_from_lengths(T, lengths):
    m = ref.Mask()(torch.zeros(len(lengths), T, 1), lengths)
    return m


# ---------------------------------------------------------------------------------------------------- ConformerBlock
def block_case(tag, D, De, stride, att, T, B, seed):
    if att.startswith("grouped"):
        att_params = {"class": "GroupedRelPosMultiHeadSelfAttention", "params": {"num_heads": 4, "group_size": int(att[7:]), "attn_drop_rate": 0.0,
                      "max_pos_encoding": 10000, "causal": False}}
    elif att == "patch":
        att_params = {"class": "RelPosPatch1dMultiHeadAttention", "params": {"num_heads": 4, "patch_size": 3, "attn_drop_rate": 0.0,
                      "num_pos_embeddings": 10000, "weight_init": "default", "bias_init": "default"}}
    else:
        att_params = {"class": "RelPos1dMultiHeadAttention", "params": {"num_heads": 4, "attn_drop_rate": 0.0,
                      "num_pos_embeddings": 10000, "weight_init": "default", "bias_init": "default"}}
    blk = ref.ConformerBlock(dim_model=D, dim_expand=De, ff_ratio=4, att_params=att_params, drop_rate=0.0, conv_stride=stride,
                             conv_params={"class": "Conv1d", "params": {"padding": "same", "kernel_size": 15}})
    blk.load_state_dict(seeded.seeded_state_dict(blk, seed))
    ref_import.zero_dropout(blk)
    blk.train()
    x = seeded.randn(tag + ".x", (B, T, D), seed).requires_grad_(True)
    lengths = torch.tensor([T] + [max(1, T - 3 - 2 * i) for i in range(B - 1)])
    mask = mask_from_lengths(T, lengths)
    y = blk(x, mask=mask)
    gy = seeded.randn(tag + ".gy", tuple(y.shape), seed)
    (y * gy).sum().backward()
    fix = {"cfg": dict(D=D, De=De, stride=stride, att=att, T=T, B=B, seed=seed), "lengths": lengths, "y": y.detach(),
           "dx": x.grad.clone(), "grads": grads_of(blk),
           "running_mean": blk.conv_module.layers[4].running_mean.clone(), "running_var": blk.conv_module.layers[4].running_var.clone()}
    blk.eval()
    with torch.no_grad():
        fix["y_eval"] = blk(x.detach(), mask=mask)
    # the reference's own bf16 envelope: same block under CPU autocast(bfloat16) vs its fp32 result
    blk.load_state_dict(seeded.seeded_state_dict(blk, seed))
    blk.train()
    xb = x.detach().clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yb = blk(xb, mask=mask)
    (yb.float() * gy).sum().backward()
    fix["env"] = {"y": rel(yb, y), "dx": rel(xb.grad, x.grad)}
    save(f"block_{tag}.pt", fix)


# -------------------------------------------------------------------------------------------------------- front-ends
def audio_frontend_case():
    pre = ref.AudioPreprocessing(sample_rate=16000, n_fft=512, win_length_ms=25, hop_length_ms=10, n_mels=80, normalize=False)
    wave = seeded.randn("wave", (3, 4000), 1, 0.1)
    wave[1, 3000:] = 0.0
    wave[2, 1700:] = 0.0
    lengths = torch.tensor([4000, 3000, 1700])
    mel, ml = pre(wave, lengths)
    save("audio_logmel.pt", {"wave_seed": 1, "lengths": lengths, "mel": mel, "mel_lengths": ml})


def resnet_block_case(tag, cin, cout, stride, hw, n, seed):
    blk = ref.ResNetBlock(in_features=cin, out_features=cout, kernel_size=(3, 3), stride=(stride, stride), act_fun="ReLU", joined_post_act=True)
    blk.load_state_dict(seeded.seeded_state_dict(blk, seed))
    blk.train()
    x = seeded.randn(tag + ".x", (n, cin, hw, hw), seed).requires_grad_(True)
    y = blk(x)
    gy = seeded.randn(tag + ".gy", tuple(y.shape), seed)
    (y * gy).sum().backward()
    fix = {"cfg": dict(cin=cin, cout=cout, stride=stride, hw=hw, n=n, seed=seed), "y": y.detach(), "dx": x.grad.clone(), "grads": grads_of(blk)}
    blk.load_state_dict(seeded.seeded_state_dict(blk, seed))
    xb = x.detach().clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yb = blk(xb)
    (yb.float() * gy).sum().backward()
    fix["env"] = {"y": rel(yb, y), "dx": rel(xb.grad, x.grad)}
    save(f"resblock_{tag}.pt", fix)


# --------------------------------------------------------------------------------------------------------- full models
def model_case(kind, seed=3):
    torch.manual_seed(0)
    if kind == "AO":
        m = ref.AudioEfficientConformerInterCTC(vocab_size=256, att_type="patch", interctc_blocks=[3, 6, 10, 13])
        B, Ls = 2, 6400
        audio = seeded.randn("audio", (B, Ls), seed, 0.1)
        alen = torch.tensor([Ls, 4800])
        audio[1, 4800:] = 0.0
        inputs = (audio, alen)
    elif kind == "VO":
        m = ref.VisualEfficientConformerInterCTC(vocab_size=256, interctc_blocks=[3, 6, 9])
        B, Tv = 2, 8
        video = seeded.randn("video", (B, Tv, 88, 88, 1), seed).clamp(-1, 1)
        vlen = torch.tensor([Tv, 6])
        video[1, 6:] = 0.0
        inputs = (video, vlen)
    else:
        m = ref.AudioVisualEfficientConformerInterCTC(vocab_size=256, v_interctc_blocks=[3, 6], a_interctc_blocks=[8, 11], f_interctc_blocks=[2])
        B, Ls, Tv = 2, 5120, 9     # Tv = Ls // 640 + 1 (align_video_to_audio, transforms.py:169-180)
        audio = seeded.randn("audio", (B, Ls), seed, 0.1)
        video = seeded.randn("video", (B, Tv, 88, 88, 1), seed).clamp(-1, 1)
        alen = torch.tensor([Ls, 3840])
        vlen = torch.tensor([Tv, 7])
        audio[1, 3840:] = 0.0
        video[1, 7:] = 0.0
        inputs = (video, vlen, audio, alen)
    m.load_state_dict(seeded.seeded_state_dict(m, seed))
    ref_import.zero_dropout(m)
    enc = m.encoder
    for e in [enc, getattr(enc, "audio_encoder", None)]:
        if e is not None and hasattr(e, "spec_augment"):
            e.spec_augment = _Bypass()
    m.train()
    outputs = m(inputs)
    labels = torch.randint(1, 256, (B, 2), generator=torch.Generator().manual_seed(seed))
    llen = torch.tensor([2, 1])
    ctc = ref.CTCLoss(zero_infinity=False, assert_shorter=False)
    keys = list(outputs.keys())
    losses = {k: ctc((labels, llen), outputs[k]) for k in keys}
    total = sum(losses.values()) / len(keys)
    total.backward()
    fix = {"kind": kind, "seed": seed, "keys": keys, "labels": labels, "label_lengths": llen,
           "logits": {k: outputs[k][0].detach() for k in keys}, "lengths": {k: outputs[k][1] for k in keys},
           "losses": {k: float(v.detach()) for k, v in losses.items()}, "total": float(total.detach()), "grads": grads_of(m, 192)}
    m.eval()
    with torch.no_grad():
        out_eval = m(inputs)
    fix["logits_eval"] = {k: out_eval[k][0] for k in keys}
    # the reference's own bf16 envelope (CPU autocast) at the logits and the loss
    m.load_state_dict(seeded.seeded_state_dict(m, seed))
    m.train()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        ob = m(inputs)
    lb = sum(float(ctc((labels, llen), [ob[k][0].float(), ob[k][1]])) for k in keys) / len(keys)
    fix["env"] = {"logits": {k: rel(ob[k][0], outputs[k][0]) for k in keys}, "total": abs(lb - float(total.detach())) / abs(float(total.detach()))}
    print(kind, "bf16 envelope", fix["env"])
    save(f"model_{kind}.pt", fix)


class _Bypass(nn.Module):
    def forward(self, x, lengths):
        return x


def state_keys():
    import json
    d = {}
    for name in ["AudioEfficientConformerInterCTC", "VisualEfficientConformerInterCTC", "AudioVisualEfficientConformerInterCTC"]:
        m = getattr(ref, name)()
        d[name] = {"params": sum(p.numel() for p in m.parameters()), "keys": {k: list(v.shape) for k, v in m.state_dict().items()}}
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(d, f)
    print("state_dict_keys.json", {k: v["params"] for k, v in d.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    state_keys()
    block_case("s1_patch_T20", 180, 180, 1, "patch", 20, 3, 11)     # T % 3 == 2: padded last patch, ragged lengths
    block_case("s1_patch_down_T21", 180, 256, 2, "patch", 21, 2, 12)  # strided block with conv_res, odd T
    block_case("s2_regular_T17", 256, 256, 1, "regular", 17, 2, 13)
    block_case("s3_regular_T9", 360, 360, 1, "regular", 9, 2, 14)
    block_case("s2_down_T12", 256, 360, 2, "regular", 12, 2, 15)
    block_case("s1_grouped3_T20", 180, 180, 1, "grouped3", 20, 3, 16)   # config-5 ablation: T % 3 == 2, d = 135
    block_case("s2_grouped1_T13", 256, 256, 1, "grouped1", 13, 2, 17)
    audio_frontend_case()
    resnet_block_case("64_64_s1", 64, 64, 1, 8, 2, 21)
    resnet_block_case("64_128_s2", 64, 128, 2, 11, 2, 22)
    resnet_block_case("128_256_s2", 128, 256, 2, 6, 3, 23)
    model_case("AO")
    model_case("VO")
    model_case("AV")
