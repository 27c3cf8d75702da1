"""CPU tests of the N > 1 path: world_size-2 gloo process groups spawned exactly like the reference's launcher
(torch.multiprocessing.spawn, main.py:188).  (i) plumbing: bucketed gradient all-reduce = mean over ranks, parameters
broadcast from rank 0; (ii) math: data parallel with LOCAL BatchNorm == independent per-shard passes + gradient mean,
checked on the pinned restatement of a ConformerBlock ('Distributed tests without a cluster')."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import seeded
from common import make_block
from oracle import restate


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from avec_b200 import parallel
        torch.manual_seed(100 + rank)
        lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        parallel.broadcast_parameters(lin)
        w0 = [p.detach().clone() for p in lin.parameters()]
        # per-rank shard
        cfg = dict(D=180, De=180, stride=1, att="patch", T=7, B=2, seed=5)
        blk, sd = make_block(cfg)
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        x = seeded.randn(f"shard{rank}", (2, 7, 180), 9)
        y, _, _ = restate.conformer_block(x, sd, torch.tensor([7, 5]), 4, 3, 1, True)
        y.square().mean().backward()
        params = [v for v in sd.values() if v.requires_grad]
        local = [p.grad.clone() for p in params]
        n = parallel.allreduce_gradients((p for p in params), world, bucket_bytes=1 << 20)   # a generator, several buckets
        reduced = [p.grad.clone() for p in params]
        # the overlapped reducer: hooks fire during backward, buckets are all-reduced as they fill up; one parameter gets no gradient
        for p in params:
            p.grad = None
        extra = torch.zeros(11, requires_grad=True)
        gb = parallel.GradientBuckets(params + [extra], bucket_bytes=1 << 20)
        y2, _, _ = restate.conformer_block(x, sd, torch.tensor([7, 5]), 4, 3, 1, True)
        y2.square().mean().backward()
        views = gb.finish()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params + [extra], views))
        torch.save({"w0": w0, "local": local, "reduced": reduced, "n": n, "overlapped": [p.grad.clone() for p in params],
                    "extra": extra.grad.clone(), "nbuckets": len(gb.buckets)}, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_broadcast_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    for a, b in zip(r[0]["w0"], r[1]["w0"]):
        assert torch.equal(a, b), "parameters were not broadcast from rank 0"
    assert r[0]["n"] == len(r[0]["local"]) > 20
    for l0, l1, g0, g1 in zip(r[0]["local"], r[1]["local"], r[0]["reduced"], r[1]["reduced"]):
        mean = (l0 + l1) / 2
        assert torch.allclose(g0, mean, rtol=1e-6, atol=1e-7) and torch.equal(g0, g1)
    assert r[0]["nbuckets"] > 2 and float(r[0]["extra"].abs().max()) == 0.0
    for g0, o0, o1 in zip(r[0]["reduced"], r[0]["overlapped"], r[1]["overlapped"]):
        assert torch.allclose(o0, g0, rtol=1e-6, atol=1e-7) and torch.equal(o0, o1)
