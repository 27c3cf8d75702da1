"""GPU parity of the training-step kernels (csrc/train.cu), all through the C ABI:
dropout masks and SpecAugment intervals BIT-EXACT against the numpy Philox restatement, ConformerBlock with dropout on
against fixtures produced by the unmodified reference fed the same masks, greedy CTC decoding bit-exact, fused Adam against
3 steps of the reference's optimizer."""
import numpy as np
import pytest
import torch

import avec_b200
import seeded
from avec_b200 import nnet, ops, functional as AF
from common import make_block, check_close, check_grad_fingerprint, rel_err, att_params
from conftest import load_golden
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu
DEV = "cuda"
SEED = 0x1234ABCD5678


def _set_rng(seed, step):
    AF.manual_seed(seed)
    ops.RNG.get(DEV).copy_(torch.tensor([seed, step], dtype=torch.int64))
    ops.RNG.site = 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,C", [(515, 180), (300, 256), (77, 30), (1, 8), (12864, 720)])
def test_dropout_mask_bit_exact(dtype, rows, C):
    _set_rng(SEED, 9)
    x = seeded.randn("drop.x", (rows, C), 1).to(DEV).to(dtype)
    res = seeded.randn("drop.r", (rows, C), 2).to(DEV).to(dtype)
    p, site = 0.1, 17
    keep = torch.from_numpy(TO.dropout_keep(SEED, 9, site, rows, C, p)).to(DEV)
    y = ops.dropout(x, p, site)
    assert torch.equal(y != 0, keep & (x != 0))
    want = torch.where(keep, x.float() * (1.0 / (1.0 - p)), torch.zeros((), device=DEV))
    check_close("y", y, want.to(dtype), 1e-6 if dtype == torch.float32 else 8e-3, 1e-7)
    y2 = ops.dropout(x, p, site, res=res, alpha=0.5)
    check_close("y+res", y2, (res.float() + 0.5 * want).to(dtype), 1e-6 if dtype == torch.float32 else 8e-3, 1e-6)
    # same (seed, step, site) -> same mask (what the backward relies on); another step -> another mask
    assert torch.equal(ops.dropout(x, p, site), y)
    ops.RNG.advance(DEV)
    keep2 = torch.from_numpy(TO.dropout_keep(SEED, 10, site, rows, C, p)).to(DEV)
    assert torch.equal(ops.dropout(x, p, site) != 0, keep2 & (x != 0))
    # in place
    xi = x.clone()
    ops.dropout(xi, p, site, out=xi)
    assert torch.equal(xi != 0, keep2 & (x != 0))


@pytest.mark.parametrize("M,N,K", [(515, 180, 720), (300, 720, 180), (1000, 256, 256), (70, 360, 1440)])
def test_dropout_fused_into_gemm_epilogue(M, N, K):
    """nn.Dropout riding in the tcgen05 GEMM's epilogue (LINEAR / SWISH / RESIDUAL forward, DSWISH backward): exactly the keep mask
    of the oracle (oracle/train_oracle.py dropout_keep) for (seed, step, site), values = mask * the undropped GEMM, and the pitched
    D = 180 rows included; the un-fused fallback (AVEC_FUSE_DROPOUT=0) agrees to bf16 rounding."""
    from avec_b200 import _lib as L
    _set_rng(SEED, 4)
    rng = ops.RNG.get(DEV)
    p, site = 0.1, 23
    x = ops.convert(seeded.randn("fdrop.x", (M, K), 1).to(DEV), torch.bfloat16, pad=True)
    w = ops.convert(0.05 * seeded.randn("fdrop.w", (N, K), 2).to(DEV), torch.bfloat16, pad=True)
    b = 0.1 * seeded.randn("fdrop.b", (N,), 3).to(DEV)
    aux = seeded.randn("fdrop.a", (M, N), 4).to(DEV).to(torch.bfloat16)
    keep = torch.from_numpy(TO.dropout_keep(SEED, 4, site, M, N, p)).to(DEV)
    scale = 1.0 / (1.0 - p)
    acc = x.float() @ w.float().t() + b
    cases = {
        "linear": (dict(epi=L.EPI_LINEAR), acc * scale, None),
        "swish": (dict(epi=L.EPI_SWISH, want_pre=True), acc * torch.sigmoid(acc) * scale, None),
        "residual": (dict(epi=L.EPI_RESIDUAL, alpha=0.5, aux=aux), 0.5 * acc * scale, aux.float()),
    }
    for name, (kw, val, res) in cases.items():
        want = torch.where(keep, val, torch.zeros((), device=DEV)) + (res if res is not None else 0.0)
        got = {}
        policy = ops.FUSE_DROPOUT
        for fused in (True, False):
            ops.FUSE_DROPOUT = 1 if fused else 0         # (1: wherever the kernel can; the default policy 2 skips wide outputs)
            try:
                n0 = avec_b200.launch_count()
                out = ops.linear_fwd(x, w, b, drop=(rng, p, site), **kw)
                got[fused] = (out, avec_b200.launch_count() - n0)
            finally:
                ops.FUSE_DROPOUT = policy
        assert got[True][1] == 1 and got[False][1] == 2, f"{name}: launches {got[True][1]} / {got[False][1]}"
        y = got[True][0][0] if name == "swish" else got[True][0]
        y0 = got[False][0][0] if name == "swish" else got[False][0]
        assert rel_err(y, want) < 4e-3, f"{name}: {rel_err(y, want)}"
        assert rel_err(y, y0) < 6e-3, f"{name} fused vs separate: {rel_err(y, y0)}"
        if res is None:
            nz = val.abs() > 1e-3                 # values that survive the bf16 rounding
            assert torch.equal((y != 0) & nz, keep & nz), name
        if name == "swish":
            assert rel_err(got[True][0][1], acc) < 4e-3       # the saved pre-activation is NOT masked
    # backward of FFN's inner dropout: drop(dswish(pre) * (dy @ W))
    dy = ops.convert(seeded.randn("fdrop.dy", (M, N), 5).to(DEV), torch.bfloat16, pad=True)
    pre = seeded.randn("fdrop.pre", (M, K), 6).to(DEV).to(torch.bfloat16)
    keep_k = torch.from_numpy(TO.dropout_keep(SEED, 4, site, M, K, p)).to(DEV)
    sg = torch.sigmoid(pre.float())
    val = (dy.float() @ w.float()) * (sg * (1 + pre.float() * (1 - sg))) * scale
    want = torch.where(keep_k, val, torch.zeros((), device=DEV))
    policy, ops.FUSE_DROPOUT = ops.FUSE_DROPOUT, 1
    try:
        n0 = avec_b200.launch_count()
        dpre = ops.linear_dgrad(dy, w, L.EPI_DSWISH, aux=pre, drop=(rng, p, site))
        assert avec_b200.launch_count() - n0 == 1
    finally:
        ops.FUSE_DROPOUT = policy
    assert rel_err(dpre, want) < 4e-3, rel_err(dpre, want)
    nz = val.abs() > 1e-3
    assert torch.equal((dpre != 0) & nz, keep_k & nz)


def test_dropout_after_patch_upsampling():
    _set_rng(SEED, 0)
    B, T, P, C = 3, 20, 3, 180
    Tp = -(-T // P)
    x = seeded.randn("dropup.x", (B * Tp, C), 1).to(DEV)
    res = seeded.randn("dropup.r", (B * T, C), 2).to(DEV)
    y = ops.dropout(x, 0.1, 5, res=res, up=(T, Tp, P))
    m = torch.from_numpy(TO.dropout_scale_mask(SEED, 0, 5, B * T, C, 0.1)).to(DEV)
    up = x.view(B, Tp, C).repeat_interleave(P, dim=1)[:, :T].reshape(B * T, C)
    check_close("y", y, res + up * m, 1e-6, 1e-6)


def test_spec_augment_matches_reference_fixture():
    fix = load_golden("train_spec_augment.pt")
    B, M, F = fix["shape"]
    _set_rng(fix["seed"], fix["step"])
    mel = seeded.randn("specaug.mel", (B, M, F), fix["mel_seed"]).transpose(1, 2).contiguous().to(DEV)
    iv = ops.spec_augment_(mel, fix["lengths"].to(DEV), fix["site"], *fix["params"], want_intervals=True)
    assert torch.equal(mel.cpu(), fix["out"].transpose(1, 2).contiguous())
    want_iv = TO.spec_augment_intervals(fix["seed"], fix["step"], fix["site"], fix["lengths"].tolist(), F, M, *fix["params"])
    assert iv.cpu().tolist() == [[list(t) for t in row] for row in want_iv]
    # module API in the reference's (B, n_mels, T) layout
    _set_rng(fix["seed"], fix["step"])
    sa = nnet.SpecAugment(*fix["params"]).train()
    out = sa(seeded.randn("specaug.mel", (B, M, F), fix["mel_seed"]).to(DEV), fix["lengths"])
    assert torch.equal(out.cpu(), fix["out"])


def test_spec_augment_full_size_properties():
    """BASELINE shape (64 x 401 x 80): idempotent, only zeroes, at most mF*F bins and mT*int(pS*len) frames per utterance"""
    _set_rng(3, 1)
    B, F, M = 64, 401, 80
    mel = (torch.randn(B, F, M, device=DEV) - 5.0)
    lengths = torch.randint(200, 402, (B,), device=DEV)
    ref = mel.clone()
    ops.spec_augment_(mel, lengths, 1)
    once = mel.clone()
    ops.spec_augment_(mel, lengths, 1)
    assert torch.equal(mel, once)
    z = mel == 0
    assert torch.equal(mel[~z], ref[~z])
    bins = z.all(dim=1).sum(dim=1)
    assert int(bins.max()) <= 2 * 27 and torch.equal(z.all(dim=1)[0], z.all(dim=1)[-1])          # shared frequency masks
    frames = z.all(dim=2).sum(dim=1)
    assert bool((frames <= 5 * (0.05 * lengths.float()).floor()).all())


def test_greedy_decode_matches_reference_fixture():
    fix = load_golden("train_greedy.pt")
    dec = nnet.CTCGreedySearchDecoder(None, blank_token=0)
    logits = fix["logits"].to(DEV)
    assert dec.greedy_search(logits, fix["lengths"]) == fix["tokens"]
    tokens, ntok, align = dec.greedy_search_device(logits, fix["lengths"], want_align=True)
    T = logits.shape[1]
    valid = torch.arange(T)[None, :] < fix["lengths"][:, None]
    assert torch.equal(align.cpu().long()[valid], fix["align"][valid]) and bool((align.cpu()[~valid] == -1).all())


def test_greedy_decode_full_size_against_torch():
    B, T, V = 64, 51, 256
    logits = torch.randn(B, T, V, device=DEV)
    logits[:, :, 0] += 2.5                     # plenty of blanks
    logits[:, 10:14, 7] += 9.0                 # repeats
    lengths = torch.randint(1, T + 1, (B,))
    got = nnet.CTCGreedySearchDecoder(None).greedy_search(logits, lengths)
    assert got == TO.greedy_decode(logits.cpu().numpy(), lengths.tolist(), 0)


def test_fused_adam_matches_reference_fixture():
    fix = load_golden("train_adam.pt")
    h = fix["hyper"]
    param = torch.nn.Parameter(fix["p0"].clone().to(DEV))
    opt = nnet.optimizers.Adam([param], lr=nnet.schedulers.NoamDecayScheduler(h["warmup_steps"], h["dim_decay"], h["val_factor"]),
                               betas=h["betas"], eps=h["eps"], weight_decay=h["weight_decay"], grad_max_norm=h["max_norm"],
                               ema_tau=h["ema_tau"])
    for i, g in enumerate(fix["grads"]):
        param.grad = g.clone().to(DEV)
        opt.step()
        info = opt.last_info()
        assert abs(info["lr"] - fix["lr"][i]) <= 1e-5 * fix["lr"][i]
        assert abs(info["grad_norm"] - fix["gnorm"][i]) <= 1e-5 * fix["gnorm"][i]
        check_close(f"p[{i}]", param, fix["p"][i], 1e-5, 1e-7)
        check_close(f"ema[{i}]", opt.ema_parameters()[0], fix["ema"][i], 1e-5, 1e-7)
    check_close("m", opt.state[param]["exp_avg"], fix["m"], 1e-5, 1e-7)
    check_close("v", opt.state[param]["exp_avg_sq"], fix["v"], 1e-5, 1e-9)
    sd = opt.state_dict()
    assert int(sd["model_step"]) == 3 and set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_fused_adam_whole_model_against_torch():
    """AO model: every parameter becomes a view of the flat buffer; one fused step == torch.optim.Adam on the same gradients"""
    torch.manual_seed(0)
    m = nnet.AudioEfficientConformerInterCTC().to(DEV)
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters()]
    grads = [torch.randn_like(p) * 0.01 for p in m.parameters()]
    opt = nnet.optimizers.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    tref = torch.optim.Adam(ref_params, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    before = avec_b200.launch_count()
    for _ in range(2):
        for p, rp, g in zip(m.parameters(), ref_params, grads):
            p.grad, rp.grad = g.clone(), g.clone()
        opt.step()
        tref.step()
    assert avec_b200.launch_count() - before == 4                       # (advance + adam) x 2 for 657 tensors
    worst = max(float((p.detach() - rp.detach()).abs().max()) for p, rp in zip(m.parameters(), ref_params))
    assert worst < 2e-6, worst
    flat = opt.flat()
    assert all(p.data_ptr() == flat["p"].data_ptr() + 4 * o for p, o in zip(flat["params"], flat["offs"]))
    # gradients written straight into the flat views need no gather
    for p, gv, g in zip(m.parameters(), opt.grad_views(), grads):
        gv.copy_(g)
        p.grad = gv
    opt.step()


@pytest.mark.parametrize("tag", ["s1_patch_T20", "s2_down_T12", "s1_grouped3_T20"])
def test_block_with_dropout_matches_reference_fixture(tag):
    fix = load_golden(f"train_block_dropout_{tag}.pt")
    cfg, p = fix["cfg"], fix["p"]
    try:
        avec_b200.set_compute_dtype(torch.float32)
        blk, sd = make_block(cfg)
        for d in blk.modules():
            if isinstance(d, torch.nn.Dropout):
                d.p = p
        blk = blk.to(DEV).train()
        x = seeded.randn(tag + ".x", (cfg["B"], cfg["T"], cfg["D"]), cfg["seed"]).to(DEV).requires_grad_(True)
        avec_b200.new_step()
        _set_rng(fix["rng_seed"], fix["step"])
        y = blk(x, klen=fix["lengths"].to(DEV).to(torch.int32))
        assert ops.RNG.site == 6
        gy = seeded.randn(tag + ".gy", tuple(y.shape), cfg["seed"]).to(DEV)
        (y * gy).sum().backward()
        check_close("y", y, fix["y"], 1e-3, 1e-4)
        check_close("dx", x.grad, fix["dx"], 1e-3, 1e-4)
        for k, fp in fix["grads"].items():
            check_grad_fingerprint(k, dict(blk.named_parameters())[k].grad, fp, 1e-3, 2e-4)
        # bf16 production mode: same masks, within the bf16 envelope of the block tests
        avec_b200.set_compute_dtype(torch.bfloat16)
        avec_b200.new_step()
        _set_rng(fix["rng_seed"], fix["step"])
        yb = blk(x.detach().to(torch.bfloat16), klen=fix["lengths"].to(DEV).to(torch.int32))
        assert rel_err(yb, fix["y"]) < 2e-2
    finally:
        avec_b200.set_compute_dtype(torch.bfloat16)


def test_model_training_graph_fresh_masks_and_determinism():
    """AO model in train() (dropout 0.1 + SpecAugment): finite loss and gradients, a new step draws new masks, the same
    (seed, step) reproduces the masks, and a captured CUDA graph draws fresh masks on every replay"""
    torch.manual_seed(0)
    m = nnet.AudioEfficientConformerInterCTC().to(DEV).train()
    m.compile(losses=nnet.CTCLoss(zero_infinity=True, assert_shorter=False))
    B = 4
    audio = (0.1 * torch.randn(B, 16000)).to(DEV)
    alen = torch.tensor([16000, 15000, 12000, 9000], device=DEV)
    labels = torch.randint(1, 256, (B, 5), device=DEV)
    lab_len = torch.full((B,), 5, device=DEV)

    def step():
        out = m((audio, alen))
        loss = m.compute_loss(out, (labels, lab_len))
        m.zero_grad(set_to_none=True)
        loss.backward()
        return out["outputs"][0].detach().clone(), float(loss)

    AF.manual_seed(77)
    o1, l1 = step()
    o2, l2 = step()
    assert np.isfinite(l1) and np.isfinite(l2) and rel_err(o2, o1) > 0.05          # new step, new masks
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    AF.manual_seed(77)
    o1b, l1b = step()
    # same (seed, step): same masks; what is left is the summation-order noise of the atomically accumulated BatchNorm
    # statistics seen through bf16 rounding
    assert rel_err(o1b, o1) < 0.3 * rel_err(o2, o1)
    m.eval()
    with torch.no_grad():
        e1 = m((audio, alen))["outputs"][0].clone()
        e2 = m((audio, alen))["outputs"][0].clone()
    assert torch.equal(e1, e2)
    m.train()
    # CUDA graph: the RNG step is advanced by a kernel inside the graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            m((audio, alen))
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_out = m((audio, alen))["outputs"][0]
    g.replay()
    r1 = static_out.clone()
    g.replay()
    r2 = static_out.clone()
    assert torch.isfinite(r1).all() and rel_err(r2, r1) > 0.05


def test_train_step_reduces_the_loss():
    """Model.compile(optimizer=...) + Model.train_step: forward + weighted CTC losses + backward + fused Adam on a fixed batch;
    the loss falls, parameters stay views of the flat buffer, compile("Adam") builds the reference's default optimizer"""
    torch.manual_seed(0)
    m = nnet.zero_dropout(nnet.AudioEfficientConformerInterCTC()).to(DEV).train()
    m.compile(losses=nnet.CTCLoss(zero_infinity=True, assert_shorter=False),
              optimizer=nnet.optimizers.Adam(m.parameters(), lr=2e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6, grad_max_norm=5.0))
    assert m.loss_weights == [0.125, 0.125, 0.125, 0.125, 0.5]
    B = 4
    audio = (0.1 * torch.randn(B, 16000)).to(DEV)
    alen = torch.full((B,), 16000, device=DEV)
    targets = (torch.randint(1, 256, (B, 5), device=DEV), torch.full((B,), 5, device=DEV))
    losses = [float(m.train_step((audio, alen), targets)) for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.8 * losses[0], losses
    flat = m.optimizer.flat()
    assert all(p.data_ptr() == flat["p"].data_ptr() + 4 * o for p, o in zip(flat["params"], flat["offs"]))
    assert int(flat["step"]) == 8 and m.optimizer.last_info()["grad_norm"] > 0
    m2 = nnet.AudioVisualEfficientConformerInterCTC()
    m2.compile(optimizer="Adam")
    o = m2.optimizer
    assert isinstance(o, nnet.optimizers.Adam) and o.scheduler.device_params() == (2 * 360 ** -0.5, 10000.0)
    assert o.param_groups[0]["betas"] == (0.9, 0.98) and o.param_groups[0]["weight_decay"] == 1e-6


def test_fused_adam_checkpoint_round_trip_with_torch_adam():
    """state_dict() is torch.optim.Adam's layout (the reference's Adam IS torch.optim.Adam + model_step, optimizers.py:61-93): a
    checkpoint written by the fused optimizer continues identically in torch's Adam and comes back; per-parameter `step` entries are
    separate tensors (a shared one would be incremented once per parameter); parameters without a gradient are left untouched."""
    torch.manual_seed(0)
    shapes = [(64, 48), (48,), (32, 64), (7,), (5, 3, 3)]
    ps_a = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    hyper = dict(betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    fused = nnet.optimizers.Adam(ps_a, lr=1e-3, **hyper)
    plain = torch.optim.Adam(ps_b, lr=1e-3, **hyper)

    def grads(step, skip=None):
        g = torch.Generator().manual_seed(100 + step)
        return [None if i == skip else torch.randn(s, generator=g).to(DEV) for i, s in enumerate(shapes)]

    for step in range(3):
        for opt_ps, opt in ((ps_a, fused), (ps_b, plain)):
            for p, g in zip(opt_ps, grads(step, skip=3 if step == 1 else None)):   # step 1: parameter 3 has no gradient
                p.grad = None if g is None else g.clone()
            opt.step()
    for i, (a, b) in enumerate(zip(ps_a, ps_b)):
        # parameter 3 skipped a step: torch keeps a per-parameter counter (2 updates), the fused kernel one device counter (its
        # bias correction is one step ahead for that parameter) - same moments, slightly different step size
        check_close(f"param[{i}]", a, b, 1e-5 if i != 3 else 1e-2, 1e-7 if i != 3 else 2e-3)
    sd = fused.state_dict()
    steps = [st["step"] for st in sd["state"].values()]
    assert all(float(s) == 3.0 and s.dtype == torch.float32 and s.dim() == 0 for s in steps)
    assert len({s.data_ptr() for s in steps}) == len(steps) and int(sd["model_step"]) == 3
    # fused -> torch Adam: one more step on both sides
    sd_plain = {k: v for k, v in sd.items() if k != "model_step"}
    ps_c = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    cont = torch.optim.Adam(ps_c, lr=1e-3, **hyper)
    cont.load_state_dict(sd_plain)
    for opt_ps, opt in ((ps_a, fused), (ps_c, cont)):
        for p, g in zip(opt_ps, grads(7)):
            p.grad = g.clone()
        opt.step()
    assert {float(st["step"]) for st in cont.state_dict()["state"].values()} == {4.0}
    # parameter 3 skipped step 1 in torch (its own counter is 3 there); the fused optimizer keeps ONE device counter, so its bias
    # correction for that parameter is one step ahead: compare the others exactly, parameter 3 loosely
    for i, (a, c) in enumerate(zip(ps_a, ps_c)):
        check_close(f"continued[{i}]", a, c, 1e-5 if i != 3 else 1e-2, 1e-7 if i != 3 else 1e-3)
    # torch Adam -> fused
    sd_back = cont.state_dict()
    sd_back["model_step"] = torch.tensor(4)
    ps_d = [torch.nn.Parameter(p.detach().clone()) for p in ps_c]
    back = nnet.optimizers.Adam(ps_d, lr=1e-3, **hyper)
    back.load_state_dict(sd_back)
    for opt_ps, opt in ((ps_c, cont), (ps_d, back)):
        for p, g in zip(opt_ps, grads(9)):
            p.grad = g.clone()
        opt.step()
    for i, (c, d) in enumerate(zip(ps_c, ps_d)):
        check_close(f"back[{i}]", c, d, 1e-5, 1e-7)
    assert int(back.state_dict()["model_step"]) == 5


def test_video_augment_matches_reference_fixture():
    """RandomCrop + RandomHorizontalFlip + TimeMaskSecond on the device: the draws bit-exact, the clip equal to the fixture the
    reference's / torchvision's own modules produced from the same draws (mean fill values to fp32 summation order)"""
    fix = load_golden("train_video_augment.pt")
    B, T, Hi, Wi = fix["shape"]
    video = seeded.randn("videoaug.x", fix["shape"], fix["video_seed"]).clamp(-1, 1)
    _set_rng(fix["seed"], fix["step"])
    out, draws = ops.video_augment(video.to(DEV), fix["lengths"].to(DEV), fix["site"], want_draws=True, rng=ops.RNG.get(DEV))
    draws = draws.cpu()
    for b in range(B):
        oy, ox, flip, masks = fix["draws"][b]
        assert draws[b, :3].tolist() == [oy, ox, int(flip)]
        assert [tuple(draws[b, 3 + 2 * m: 5 + 2 * m].tolist()) for m in range(len(masks))] == [tuple(iv) for iv in masks]
    want = TO.video_augment(video.numpy(), fix["lengths"].tolist(), fix["seed"], fix["step"], fix["site"])
    assert float((out.cpu() - torch.from_numpy(want)).abs().max()) <= 1e-6
    assert float((out.cpu()[:, ::3, ::4, ::4] - fix["out_sub"]).abs().max()) <= 1e-6
    # module interface: train() augments, eval() is the configs' CenterCrop; frames past the length are zero
    aug = nnet.VideoAugment().to(DEV).train()
    y = aug(video.to(DEV).unsqueeze(-1), fix["lengths"].to(DEV))
    assert y.shape == (B, T, 88, 88, 1) and float(y[2, int(fix["lengths"][2]):].abs().max()) == 0.0
    ye = aug.eval()(video.to(DEV), None)
    assert torch.equal(ye, video.to(DEV)[:, :, 4:92, 4:92])
