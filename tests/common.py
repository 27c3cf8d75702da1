"""helpers shared by the parity tests"""
import torch

import seeded


def att_params(att):
    base = {"num_heads": 4, "attn_drop_rate": 0.0, "num_pos_embeddings": 10000, "weight_init": "default", "bias_init": "default"}
    if att.startswith("grouped"):
        return {"class": "GroupedRelPosMultiHeadSelfAttention", "params": {"num_heads": 4, "group_size": int(att[7:]), "attn_drop_rate": 0.0,
                                                                            "max_pos_encoding": 10000, "causal": False}}
    if att == "patch":
        return {"class": "RelPosPatch1dMultiHeadAttention", "params": dict(base, patch_size=3)}
    return {"class": "RelPos1dMultiHeadAttention", "params": base}


def make_block(cfg):
    from avec_b200 import nnet
    blk = nnet.ConformerBlock(dim_model=cfg["D"], dim_expand=cfg["De"], ff_ratio=4, att_params=att_params(cfg["att"]), drop_rate=0.0,
                              conv_stride=cfg["stride"], conv_params={"class": "Conv1d", "params": {"padding": "same", "kernel_size": 15}})
    sd = seeded.seeded_state_dict(blk, cfg["seed"])
    blk.load_state_dict(sd)
    return blk, sd


def block_G(cfg):
    return int(cfg["att"][7:]) if cfg["att"].startswith("grouped") else None


def block_lengths(cfg):
    T, B = cfg["T"], cfg["B"]
    return torch.tensor([T] + [max(1, T - 3 - 2 * i) for i in range(B - 1)])


def check_close(name, got, want, rtol, atol):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = (err > tol)
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{name}: {int(bad.sum())}/{bad.numel()} elements out of tolerance (rtol={rtol}, atol={atol}); "
                             f"max abs err {float(err.max()):.3e} at flat index {i}: got {float(got.flatten()[i]):.6e} want "
                             f"{float(want.flatten()[i]):.6e}; ref absmax {float(want.abs().max()):.3e}")


def check_grad_fingerprint(name, grad, fp, rtol, atol):
    f = grad.detach().float().cpu().flatten()
    check_close(name, f[::fp["stride"]], fp["vals"], rtol, atol)


def rel_err(got, want):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    return float((got - want).norm() / (want.norm() + 1e-30))
