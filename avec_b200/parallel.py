"""Data-parallel plumbing: one process per GPU, full replica per rank, batch axis sharded, local
BatchNorm statistics, ONE collective per step - the gradient all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
Replaces the reference's DDP wrap (nnet/model.py:59-65) on the hot path."""
import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """identical replicas: parameters and buffers of rank `src` to every rank (DDP does this at construction)"""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def allreduce_gradients(params, world_size=None, bucket_bytes=256 << 20):
    """mean of the gradients over ranks, in a few large flat buckets (61.7 M parameters = 247 MB fp32 for AV)."""
    world_size = world_size or dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    for g in grads + [None]:
        if g is not None and (size + g.numel() * g.element_size() <= bucket_bytes or not bucket):
            bucket.append(g)
            size += g.numel() * g.element_size()
            continue
        if bucket:
            flat = torch._utils._flatten_dense_tensors(bucket)
            dist.all_reduce(flat)
            flat.div_(world_size)
            for dst, src in zip(bucket, torch._utils._unflatten_dense_tensors(flat, bucket)):
                dst.copy_(src)
        bucket, size = ([g], g.numel() * g.element_size()) if g is not None else ([], 0)
    return len(grads)
