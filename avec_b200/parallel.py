"""Data-parallel plumbing: one process per GPU, full replica per rank, batch axis sharded, local
BatchNorm statistics, ONE collective per step - the gradient all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
Replaces the reference's DDP wrap (nnet/model.py:59-65) on the hot path."""
import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """identical replicas: parameters and buffers of rank `src` to every rank (DDP does this at construction)"""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def allreduce_gradients(params, world_size=None, bucket_bytes=256 << 20, grads=None):
    """mean of the gradients over ranks in a few large flat buckets (61.7 M parameters = 247 MB fp32 for AV).

    Per bucket: ONE flatten (torch.cat), one all-reduce (ReduceOp.AVG on NCCL; sum + scale on gloo), and the averaged values
    are handed back as VIEWS of the flat buffer (`p.grad = view`) - no per-tensor copy kernels (1100 launches per step
    otherwise).  `grads` (optional) are the tensors to reduce when they are not `p.grad` itself, e.g. the static gradient
    tensors a captured CUDA graph writes on every replay."""
    world_size = world_size or dist.get_world_size()
    params = list(params)          # `params` may be a generator (model.parameters()): walk it exactly once
    if grads is None:
        pairs = [(p, p.grad) for p in params]
    else:
        grads = list(grads)
        if len(grads) != len(params):
            raise ValueError(f"allreduce_gradients: {len(grads)} gradients for {len(params)} parameters")
        pairs = list(zip(params, grads))
    pairs = [(p, g) for p, g in pairs if g is not None]
    avg = dist.get_backend() == "nccl"
    bucket, size = [], 0
    for item in pairs + [None]:
        nbytes = item[1].numel() * item[1].element_size() if item is not None else 0
        if item is not None and (size + nbytes <= bucket_bytes or not bucket):
            bucket.append(item)
            size += nbytes
            continue
        if bucket:
            flat = torch.cat([g.reshape(-1) for _, g in bucket])
            if avg:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(flat)
                flat.div_(world_size)
            off = 0
            for p, g in bucket:
                n = g.numel()
                p.grad = flat[off:off + n].view_as(g)
                off += n
        bucket, size = ([item], nbytes) if item is not None else ([], 0)
    return len(pairs)
