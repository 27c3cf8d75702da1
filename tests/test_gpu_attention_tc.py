"""tcgen05 / TMEM / TMA relative-position attention (csrc/attention_tc.cu) against a plain PyTorch fp32 statement of
RelPos1dMultiHeadAttention.forwardQKV + rel_to_abs (reference nnet/attentions.py:280-323, 258-276) on the same bf16-rounded
operands: forward output and log-sum-exp, and the backward (dq, dk, dv, de) against torch autograd, for one-tile heads
(T <= 128: every BASELINE shape), multi-tile sequences (flash loop, fp32 partial-sum accumulation), ragged key lengths and the
fully masked query rows of a zero-padded last patch.  Tolerance: relative L2 2e-2 (bf16 P / dS operands), stated below."""
import pytest
import torch

from avec_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference(q, k, v, e, klen, qlen):
    """q, k, v [B,H,T,d], e [H,2T-1,d] fp32 -> o [B,H,T,d], lse [B,H,T]  (scores masked with -1e9 as the reference does)"""
    B, H, T, d = q.shape
    s_k = q @ k.transpose(2, 3)
    s_rel = q @ e.transpose(1, 2).unsqueeze(0)                                        # [B,H,T,2T-1]
    idx = (T - 1) + torch.arange(T, device=q.device)[None, :] - torch.arange(T, device=q.device)[:, None]
    s = (s_k + s_rel.gather(3, idx.expand(B, H, T, T))) / d ** 0.5
    keep = (torch.arange(T, device=q.device)[None, None, None, :] < klen[:, None, None, None]).float()
    keep = keep * (torch.arange(T, device=q.device)[None, None, :, None] < qlen).float()
    s = s + (1.0 - keep) * -1e9
    return s.softmax(-1) @ v, torch.logsumexp(s, -1)


def _pad_heads(t, H, d, dp):
    """[rows, n*H*d] -> [rows, n*H*dp]"""
    rows = t.shape[0]
    out = torch.zeros(rows, t.shape[1] // d, dp, device=t.device, dtype=t.dtype)
    out[:, :, :d] = t.view(rows, -1, d)
    return out.view(rows, -1)


def _unpad_heads(t, H, d, dp):
    rows = t.shape[0]
    return t.view(rows, -1, dp)[:, :, :d].reshape(rows, -1)


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


@pytest.mark.parametrize("B,T,H,d,ragged,qlen_cut", [
    (3, 67, 4, 45, True, 0),      # audio stage 1 (patch attention on 201 / 3 tokens)
    (3, 101, 4, 64, True, 0),     # stage 2 / video
    (2, 51, 4, 90, True, 0),      # stage 3 / fusion
    (2, 7, 4, 45, True, 1),       # zero-padded last patch: the last query row sees every key masked
    (2, 128, 4, 64, False, 0),    # exactly one tile
    (2, 129, 4, 64, True, 0),     # one key past the tile
    (2, 200, 4, 90, True, 0),     # multi-tile, two head-dim blocks
    (1, 400, 4, 64, True, 0),     # long sequence (ablation sweep, stage 2 at 1600 frames)
    (1, 267, 4, 45, True, 1),     # patch attention at 1600 frames with a masked last row
])
def test_attention_tc_matches_torch(B, T, H, d, ragged, qlen_cut):
    torch.manual_seed(T * 7 + d)
    dp = ops.attn_head_pad(d)
    D = H * d
    qkv = (0.7 * torch.randn(B * T, 3 * D, device=DEV)).to(torch.bfloat16)
    e = (0.7 * torch.randn(2 * T - 1, D, device=DEV)).to(torch.bfloat16)
    klen = torch.tensor([T if (b == 0 or not ragged) else max(1, T - 5 - 9 * b) for b in range(B)], dtype=torch.int32, device=DEV)
    qlen = T - qlen_cut
    go = (0.5 * torch.randn(B * T, D, device=DEV)).to(torch.bfloat16)

    qkv_p, e_p, go_p = _pad_heads(qkv, H, d, dp), _pad_heads(e, H, d, dp), _pad_heads(go, H, d, dp)
    o_p, lse = ops.relpos_attn_tc_fwd(qkv_p, e_p, klen, qlen, B, T, H, d, dp)
    dqkv_p, de_p = ops.relpos_attn_tc_bwd(go_p, qkv_p, e_p, o_p, lse, klen, qlen, B, T, H, d, dp)
    torch.cuda.synchronize()

    qf = qkv.float().view(B, T, 3, H, d).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)     # [3,B,H,T,d]
    ef = e.float().view(2 * T - 1, H, d).transpose(0, 1).contiguous().requires_grad_(True)
    o_ref, lse_ref = _reference(qf[0], qf[1], qf[2], ef, klen.long(), qlen)
    o_ref2 = o_ref.transpose(1, 2).reshape(B * T, D)
    (o_ref2 * go.float()).sum().backward()
    dqkv_ref = qf.grad.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * D)
    de_ref = ef.grad.transpose(0, 1).reshape(2 * T - 1, D)

    o = _unpad_heads(o_p, H, d, dp)
    assert float(o_p.view(B * T, H, dp)[:, :, d:].abs().max()) == 0.0 if dp > d else True       # pad columns stay exact zeros
    assert _rel(o, o_ref2) < 1e-2, f"o rel L2 {_rel(o, o_ref2)}"
    # rows whose every key is masked: the reference's lse is log(T) - 1e9 (all scores shifted); compare the others
    rows_ok = torch.arange(T, device=DEV) < qlen
    assert float((lse[:, :, rows_ok] - lse_ref[:, :, rows_ok]).abs().max()) < 2e-2
    dqkv = _unpad_heads(dqkv_p, H, d, dp)
    de = _unpad_heads(de_p, H, d, dp)
    for name, got, want in (("dq", dqkv[:, :D], dqkv_ref[:, :D]), ("dk", dqkv[:, D:2 * D], dqkv_ref[:, D:2 * D]),
                            ("dv", dqkv[:, 2 * D:], dqkv_ref[:, 2 * D:]), ("de", de, de_ref)):
        assert _rel(got, want) < 2e-2, f"{name} rel L2 {_rel(got, want)} (T={T}, d={d})"
    if dp > d:
        assert float(dqkv_p.view(B * T, 3 * H, dp)[:, :, d:].abs().max()) == 0.0


@pytest.mark.parametrize("G,T,D", [(3, 20, 180), (3, 401, 180), (1, 13, 256), (1, 200, 360)])
def test_grouped_attention_tc_matches_simt_kernels(G, T, D):
    """GroupedRelPosMultiHeadSelfAttention (Transformer-XL biases u, v; tokens of G frames, T not a multiple of G) through the tile
    kernel (frames regrouped into the 4-part padded-heads layout) against the round-1 SIMT kernels on the same bf16 module:
    output, input gradient and every parameter gradient incl. du / dv"""
    import avec_b200
    from avec_b200 import nnet
    avec_b200.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(G * 100 + T)
    att = {"class": "GroupedRelPosMultiHeadSelfAttention", "params": {"num_heads": 4, "group_size": G, "attn_drop_rate": 0.0, "max_pos_encoding": 10000,
                                                                         "causal": False}}
    mod = nnet.AttentionModule(D, att, 0.0).to(DEV).train()
    with torch.no_grad():
        mod.attention.u.normal_(0, 0.3)
        mod.attention.v.normal_(0, 0.3)
        for lin in (mod.attention.query_layer, mod.attention.key_layer, mod.attention.pos_layer):
            lin.weight.mul_(3.0)
    B = 2
    x = torch.randn(B, T, D, device=DEV).to(torch.bfloat16)
    klen = torch.tensor([T, max(1, T - 7)], dtype=torch.int32, device=DEV)
    gy = torch.randn(B, T, D, device=DEV)
    res = {}
    for tc in (True, False):
        ops.set_attention_tc(tc)
        try:
            xi = x.clone().requires_grad_(True)
            mod.zero_grad(set_to_none=True)
            y = mod.forward_residual(xi, klen)
            (y.float() * gy).sum().backward()
            res[tc] = (y.detach(), xi.grad.detach(), {k: p.grad.detach().clone() for k, p in mod.named_parameters()})
        finally:
            ops.set_attention_tc(True)
    assert _rel(res[True][0], res[False][0]) < 1e-2 and _rel(res[True][1], res[False][1]) < 2e-2
    for k, g in res[False][2].items():
        if k.endswith("key_layer.bias") or k.endswith("pos_layer.bias"):
            continue        # analytically zero (softmax is invariant to a per-row constant): rounding noise on both sides
        assert _rel(res[True][2][k], g) < 3e-2, f"{k}: {_rel(res[True][2][k], g)}"
