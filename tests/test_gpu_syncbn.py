"""2-GPU test of the SyncBatchNorm switch (ops.set_sync_batchnorm; the reference's DDP default, nnet/model.py:59-61):
two ranks with half the batch each and synchronised statistics must reproduce the single-process full-batch result
(outputs, input gradients, summed parameter gradients, running statistics) of a ConformerBlock and a strided ResNetBlock.
Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(kind):
    import seeded
    from avec_b200 import nnet
    from common import make_block
    if kind == "conformer":
        cfg = dict(D=256, De=360, stride=2, att="regular", T=12, B=4, seed=31)
        blk, _ = make_block(cfg)
        x = seeded.randn("sync.x", (4, 12, 256), 31)
    else:
        blk = nnet.ResNetBlock(64, 128, (3, 3), (2, 2), "ReLU", True)
        blk.load_state_dict(seeded.seeded_state_dict(blk, 33))
        x = seeded.randn("sync.img", (4, 12, 12, 64), 33)
    return blk, x


def _run(blk, x, kind):
    import avec_b200
    avec_b200.new_step()
    x = x.clone().requires_grad_(True)
    y = blk(x, klen=None) if kind == "conformer" else blk(x)
    g = torch.linspace(-1, 1, y[0].numel(), device=y.device).view(y.shape[1:])
    (y * g).sum().backward()
    return y.detach(), x.grad.detach(), [p.grad.detach().clone() for p in blk.parameters()]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import avec_b200
        from avec_b200 import ops
        avec_b200.set_compute_dtype(torch.float32)
        res = {}
        for kind in ("conformer", "resnet"):
            blk, x = _build(kind)
            blk = blk.to(dev).train()
            state0 = {k: v.clone() for k, v in blk.state_dict().items()}
            ops.set_sync_batchnorm(False)
            yf, dxf, gf = _run(blk, x.to(dev), kind)                           # full batch, local statistics
            bufs_full = {k: v.clone() for k, v in blk.named_buffers()}
            blk.load_state_dict(state0)
            blk.zero_grad(set_to_none=True)
            ops.set_sync_batchnorm(True)
            sl = slice(2 * rank, 2 * rank + 2)
            ys, dxs, gs = _run(blk, x[sl].to(dev), kind)                       # half batch, synchronised statistics
            ops.set_sync_batchnorm(False)
            for g in gs:
                dist.all_reduce(g)
            err = {"y": float((ys - yf[sl]).abs().max()), "dx": float((dxs - dxf[sl]).abs().max()),
                   # (+1e-3: the bias gradients in front of a BatchNorm are exactly 0 in exact arithmetic and pure rounding noise here)
                   "grads": max(float((a - b).abs().max() / (b.abs().max() + 1e-3)) for a, b in zip(gs, gf)),
                   "buffers": max(float((v.float() - bufs_full[k].float()).abs().max()) for k, v in blk.named_buffers()),
                   "scale": float(yf.abs().max())}
            res[kind] = err
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def test_sync_batchnorm_two_ranks_equal_full_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    for kind, e in res.items():
        assert e["y"] < 2e-4 * max(1.0, e["scale"]), (kind, e)
        assert e["dx"] < 2e-4 * max(1.0, e["scale"]), (kind, e)
        # parameter gradients are fp32 atomic sums over the batch (order varies run to run): the analytically-zero ones (key / position
        # biases, biases in front of a BatchNorm) sit at the 1e-6 noise floor, which the +1e-3 denominator turns into a few 1e-3
        assert e["grads"] < 5e-3, (kind, e)
        assert e["buffers"] < 1e-5, (kind, e)
