"""Parity AT THE BENCHMARKED CONFIGURATION, in the benchmarked precision (VERDICT r01 'what's weak' 1): the bf16 tcgen05 /
tensor-core production path at BASELINE.json shapes - 4 s of audio + 101 video frames per utterance, ResNet trunk on B*T = 6464
images, stage-1 Conformer tokens M = 12864 - against the fp32 restatement (oracle/restate.py, pinned to fixtures of the unmodified
reference in the CPU suite) evaluated on the same GPU.  The reference's OWN bf16 error at these shapes is measured next to it
(the same restatement under torch.autocast(bf16) vs its fp32) and printed beside ours: run with `-s` to see the table."""
import pytest
import torch
import torch.nn.functional as F

import avec_b200
import seeded
from avec_b200 import nnet
from common import make_block, rel_err
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ctc_total(outs, labels, llen):
    tot = 0.0
    for v in outs.values():
        logp = F.log_softmax(v[0].float(), dim=-1).transpose(0, 1)
        tot = tot + F.ctc_loss(logp, labels, v[1].to(torch.long), llen, blank=0, reduction="none", zero_infinity=True).mean()
    return tot / len(outs)


def test_av_model_benchmark_shape_bf16_vs_oracle():
    """AV EffConfInterCTC, 64000 samples + 101 frames, B = 16 utterances (the largest the test budget allows; BatchNorm sees
    16 x 101 x 22 x 22 sites, the trunk runs on 1616 images), train-mode BatchNorm, dropout 0: bf16 production path vs fp32 oracle.
    Tolerances: logits relative L2 within 2x the reference graph's own bf16-autocast error at the same shape (+1e-2), total CTC
    loss within that envelope; the greedy alignment agreement is reported beside the autocast-vs-fp32 figure."""
    B, Ls, Tv = 16, 64000, 101
    g = torch.Generator().manual_seed(1234)
    audio = (0.1 * torch.randn(B, Ls, generator=g)).to(DEV)
    video = torch.randn(B, Tv, 88, 88, 1, generator=g).clamp_(-1, 1).to(DEV)
    alen = torch.tensor([Ls - 1280 * (i % 4) for i in range(B)], device=DEV)      # ragged: 51 / 49 / 47 / 45 output frames
    vlen = alen // 640 + 1
    labels = torch.randint(1, 256, (B, 20), generator=g).to(DEV)
    llen = torch.full((B,), 20, device=DEV)
    avec_b200.set_compute_dtype(torch.bfloat16)
    m = nnet.AudioVisualEfficientConformerInterCTC(vocab_size=256)
    m.load_state_dict(seeded.seeded_state_dict(m, 11))
    nnet.zero_dropout(m)
    m = m.to(DEV).train()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out = m((video, vlen, audio, alen))
    ctc = nnet.CTCLoss(zero_infinity=True, assert_shorter=False)
    loss = sum(ctc((labels, llen), v) for v in out.values()) / len(out)
    loss.backward()
    with torch.no_grad():
        want = restate.av_model(sd, video, vlen, audio, alen)
        want_loss = float(_ctc_total(want, labels, llen))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            env_out = restate.av_model(sd, video, vlen, audio, alen)
        env_loss = float(_ctc_total(env_out, labels, llen))
    rows = []
    for k in want:
        assert out[k][1].tolist() == want[k][1].tolist(), f"lengths of {k}"
        w = want[k][0].float()
        e_ours, e_env = rel_err(out[k][0], w), rel_err(env_out[k][0], w)
        valid = torch.arange(w.shape[1], device=DEV)[None, :] < want[k][1][:, None]
        ag_ours = float((out[k][0].argmax(-1) == w.argmax(-1))[valid].float().mean())
        ag_env = float((env_out[k][0].float().argmax(-1) == w.argmax(-1))[valid].float().mean())
        top2 = w.topk(2, dim=-1).values
        clear = valid & ((top2[..., 0] - top2[..., 1]) > 0.1)
        ag_clear = float((out[k][0].argmax(-1) == w.argmax(-1))[clear].float().mean()) if bool(clear.any()) else 1.0
        ag_clear_env = float((env_out[k][0].float().argmax(-1) == w.argmax(-1))[clear].float().mean()) if bool(clear.any()) else 1.0
        rows.append((k, e_ours, e_env, ag_ours, ag_env, ag_clear, ag_clear_env))
    print("\nAV @ 4 s + 101 frames, B=16, bf16 vs fp32 oracle (ours | reference graph under bf16 autocast):")
    for k, e1, e2, a1, a2, ac, ace in rows:
        print(f"  {k:10s} logits rel-L2 {e1:.4f} | {e2:.4f}   greedy agreement {100 * a1:.2f}% | {100 * a2:.2f}%   (margin > 0.1: {100 * ac:.2f}% | {100 * ace:.2f}%)")
    print(f"  total CTC loss {float(loss):.5f} | {env_loss:.5f}   oracle {want_loss:.5f}")
    for k, e_ours, e_env, ag_ours, ag_env, ag_clear, ag_clear_env in rows:
        assert e_ours < 2 * e_env + 1e-2, f"logits[{k}]: rel L2 {e_ours:.4f} vs the reference graph's own bf16 error {e_env:.4f}"
        # random-init logits are near-ties: greedy indices are compared where the oracle's top-2 margin
        # exceeds 0.1, against the agreement the reference graph itself reaches in bf16 on the same frames
        assert ag_clear >= ag_clear_env - 0.02 and ag_ours >= ag_env - 0.03, f"greedy CTC indices on {k}: {ag_clear:.4f} vs {ag_clear_env:.4f}"
    env_rel = abs(env_loss - want_loss) / abs(want_loss)
    assert abs(float(loss) - want_loss) / abs(want_loss) <= 2 * env_rel + 1e-2
    grads = [p.grad for p in m.parameters()]
    assert all(gr is not None and bool(torch.isfinite(gr).all()) for gr in grads)


def test_resnet_trunk_benchmark_shape_bf16_vs_oracle():
    """ResNet-18 trunk (8 BasicBlocks + GlobalAvgPool + Linear 512->256) on N = 6464 images of 22 x 22 x 64 = the B = 64 x 101-frame
    tensor of the benchmark, bf16 tcgen05 path (halo / whole-image / parity-class conv kernels) vs the fp32 restatement: output,
    input gradient and parameter gradients by relative L2 against the restatement's own bf16-autocast error."""
    N, H, W, C = 6464, 22, 22, 64
    avec_b200.set_compute_dtype(torch.bfloat16)
    trunk = nnet.ResNet(include_stem=False, dim_output=256, model="ResNet18")
    trunk.load_state_dict(seeded.seeded_state_dict(trunk, 5))
    trunk = trunk.to(DEV).train()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in trunk.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn(N, H, W, C, generator=g)).to(DEV)           # post-ReLU / max-pool statistics
    gy = torch.randn(N, 256, generator=g).to(DEV)
    xb = x.to(torch.bfloat16).requires_grad_(True)
    with avec_b200.functional.forward_scope(trunk, xb.device):
        y = trunk(xb)
    (y.float() * gy).sum().backward()

    def oracle(xin):
        h = xin.permute(0, 3, 1, 2)
        strides = [1, 1, 2, 1, 2, 1, 2, 1]
        for k in range(8):
            h = restate.resnet_block(h, sd, strides[k], True, f"blocks.{k}.")
        return F.linear(h.mean((2, 3)), sd["head.1.weight"], sd["head.1.bias"])

    xr = x.clone().requires_grad_(True)
    yr = oracle(xr)
    (yr * gy).sum().backward()
    ref_grads = {k: v.grad.clone() for k, v in sd.items() if v.grad is not None}
    for v in sd.values():
        v.grad = None
    xe = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ye = oracle(xe)
    (ye.float() * gy).sum().backward()
    e_y, env_y = rel_err(y, yr), rel_err(ye, yr)
    e_dx, env_dx = rel_err(xb.grad, xr.grad), rel_err(xe.grad, xr.grad)
    print(f"\nResNet trunk N=6464: y rel-L2 {e_y:.4f} (autocast {env_y:.4f}); dx rel-L2 {e_dx:.4f} (autocast {env_dx:.4f})")
    assert e_y < 2 * env_y + 5e-3 and e_dx < 2 * env_dx + 1e-2
    params = dict(trunk.named_parameters())
    worst = 0.0
    for k in ("blocks.0.layers.0.weight", "blocks.1.layers.3.weight", "blocks.2.layers.0.weight", "blocks.2.residual.0.weight",
              "blocks.3.layers.3.weight", "blocks.4.layers.0.weight", "blocks.5.layers.3.weight", "blocks.6.layers.0.weight",
              "blocks.7.layers.3.weight", "blocks.0.layers.1.weight", "blocks.4.layers.4.bias", "head.1.weight"):
        e, env = rel_err(params[k].grad, ref_grads[k]), rel_err(sd[k].grad, ref_grads[k])
        worst = max(worst, e)
        assert e < 2 * env + 2e-2, f"grad {k}: rel L2 {e:.4f} vs autocast {env:.4f}"
    print(f"  worst sampled parameter-gradient rel-L2 {worst:.4f}")


def test_stage1_block_benchmark_shape_bf16_vs_oracle():
    """one stage-1 audio ConformerBlock (D = 180, patch attention P = 3) on M = 64 x 201 = 12864 tokens in bf16 (tcgen05 GEMMs,
    tensor-core attention) vs the fp32 restatement, fwd + input gradient"""
    cfg = dict(D=180, De=180, stride=1, att="patch", T=201, B=64, seed=31)
    avec_b200.set_compute_dtype(torch.bfloat16)
    blk, sd = make_block(cfg)
    blk = blk.to(DEV).train()
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    x = seeded.randn("bench.block.x", (cfg["B"], cfg["T"], cfg["D"]), 31).to(DEV)
    gy = seeded.randn("bench.block.gy", (cfg["B"], cfg["T"], cfg["De"]), 31).to(DEV)
    klen = torch.full((cfg["B"],), cfg["T"], dtype=torch.int32, device=DEV)
    klen[1::2] -= 17
    xb = x.to(torch.bfloat16).requires_grad_(True)
    y = blk(xb, klen=klen)
    (y.float() * gy).sum().backward()
    xr = x.clone().requires_grad_(True)
    yr, _, _ = restate.conformer_block(xr, sdg, klen, 4, 3, 1, training=True)
    (yr * gy).sum().backward()
    xe = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ye, _, _ = restate.conformer_block(xe, sdg, klen, 4, 3, 1, training=True)
    (ye.float() * gy).sum().backward()
    e_y, env_y, e_dx, env_dx = rel_err(y, yr), rel_err(ye, yr), rel_err(xb.grad, xr.grad), rel_err(xe.grad, xr.grad)
    print(f"\nstage-1 block M=12864: y rel-L2 {e_y:.4f} (autocast {env_y:.4f}); dx rel-L2 {e_dx:.4f} (autocast {env_dx:.4f})")
    assert e_y < 2 * env_y + 5e-3 and e_dx < 2 * env_dx + 1e-2
