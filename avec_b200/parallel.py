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


class GradientBuckets:
    """Gradient all-reduce overlapped with the backward (the role of DDP's bucketed reducer in the reference, nnet/model.py:59-65).

    Parameters are grouped, in reverse registration order (the order their gradients become ready), into flat fp32 buckets of
    `bucket_bytes`.  A post-accumulate hook on every parameter counts arrivals; when a bucket is complete its gradients are copied
    into the flat buffer with ONE multi-tensor launch and the bucket's all-reduce is issued on a dedicated communication stream
    while the backward keeps running on the compute stream(s).  `finish()` joins the communication stream and re-points every
    `p.grad` at its (averaged) bucket view - the fused optimizer can take the flat buffer as is.  Everything (hooks, event waits,
    NCCL calls) is capturable in a CUDA graph: under capture the buckets become parallel branches of the graph.
    With `flat` (a preallocated fp32 buffer, e.g. the fused Adam's gradient buffer, and `offsets`) the buckets are slices of it.
    Measured on 2 x B200 (AV, per-GPU batch 64, profiles/r02_multigpu.md): ONE bucket reduced right after the backward costs
    +0.8 ms per step (47.1 vs 46.3 ms on one GPU); 32 MB buckets overlapped with the backward cost +4.0 ms, because the NCCL
    kernels occupy SMs that the persistent one-CTA-per-SM conv / GEMM kernels are sized for - bench.py therefore uses one bucket."""

    def __init__(self, params, bucket_bytes=32 << 20, group=None, flat=None, offsets=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        dev = self.params[0].device
        self.cuda = dev.type == "cuda"
        if offsets is None:
            offsets, n = [], 0
            for p in self.params:
                offsets.append(n)
                n += (p.numel() + 3) // 4 * 4
            total = n
        else:
            total = flat.numel()
        self.flat = flat if flat is not None else torch.zeros(total, device=dev, dtype=torch.float32)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for p, o in zip(self.params, offsets)]
        # buckets: contiguous runs of parameters, built from the LAST parameter backwards
        self.buckets, cur, size = [], [], 0
        for i in reversed(range(len(self.params))):
            nbytes = self.params[i].numel() * 4
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(i)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        self.spans = []
        for b, idxs in enumerate(self.buckets):
            lo = min(offsets[i] for i in idxs)
            hi = max(offsets[i] + (self.params[i].numel() + 3) // 4 * 4 for i in idxs)
            self.spans.append((lo, min(hi, total)))
            for i in idxs:
                self.bucket_of[i] = b
        self.comm = torch.cuda.Stream(device=dev) if self.cuda else None
        self.pending = [0] * len(self.buckets)
        self.streams = [set() for _ in self.buckets]
        self.done = [False] * len(self.buckets)
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(param):
            b = self.bucket_of[i]
            self.pending[b] += 1
            if self.cuda:
                self.streams[b].add(torch.cuda.current_stream(param.device))
            if self.pending[b] == len(self.buckets[b]):
                self._launch(b)
        return hook

    def _launch(self, b):
        idxs = self.buckets[b]
        src = [self.params[i].grad for i in idxs]
        dst = [self.views[i] for i in idxs]
        pairs = [(d, s) for d, s in zip(dst, src) if s is not None and s.data_ptr() != d.data_ptr()]
        lo, hi = self.spans[b]
        buf = self.flat[lo:hi]
        if self.cuda:
            from . import functional as AF
            AF.join_side_streams(self.flat.device)       # weight gradients computed on the side stream
            cur = torch.cuda.current_stream(self.flat.device)
            for s in self.streams[b]:
                if s != cur:
                    cur.wait_stream(s)              # gradients of this bucket produced on the other compute stream
            if pairs:
                torch._foreach_copy_([d for d, _ in pairs], [s for _, s in pairs])
            self.comm.wait_stream(cur)
            if self.world > 1:
                with torch.cuda.stream(self.comm):
                    dist.all_reduce(buf, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, group=self.group)
                    if not self.avg:
                        buf.div_(self.world)
        else:
            if pairs:
                torch._foreach_copy_([d for d, _ in pairs], [s for _, s in pairs])
            if self.world > 1:
                dist.all_reduce(buf, group=self.group)
                buf.div_(self.world)
        self.done[b] = True

    def finish(self):
        """end of backward: joins the communication stream and hands the averaged views back as p.grad.  A bucket that never
        filled up (a parameter without a gradient this step) is flushed here with zeros for the missing gradients, so every rank
        issues the same sequence of collectives."""
        if self.cuda:
            from . import functional as AF
            AF.join_side_streams(self.flat.device)
        for b in range(len(self.buckets)):
            if not self.done[b]:
                for i in self.buckets[b]:
                    if self.params[i].grad is None:
                        self.views[i].zero_()
                self._launch(b)
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm)
        for p, v in zip(self.params, self.views):
            p.grad = v
        self.pending = [0] * len(self.buckets)
        self.streams = [set() for _ in self.buckets]
        self.done = [False] * len(self.buckets)
        return self.views
