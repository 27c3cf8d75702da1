"""Kernel-layout copies of the fp32 master parameters, rebuilt by ONE launch per training step.

Every GEMM / convolution of the hot path reads its weight in a layout of its own: compute dtype (bf16), rows pitched for TMA,
conv filters permuted to [Co][tap][Ci] (forward) / [Ci][tap][Co] (dgrad), Q/K/V stacked with every head zero-padded to its column
block, the 7200 -> 180 projection in (f, c) feature order ...  A `layout(...)` call describes such a copy as strided segments of the
parameter tensors; the plan keeps ONE persistent destination buffer per layout and a device-side job table, and
`avec_convert_multi` refreshes all of them in a single launch the first time a weight is needed after the parameters changed
(functional.invalidate_weights(): fused optimizer step, checkpoint load, start of a benchmark step).  A parameter modified through
torch (its version counter / storage pointer moved) refreshes just its own layouts.

Replaces the per-weight `.to(bf16)` / permute / cat kernels autocast and cuDNN issue inside every reference forward
(nnet/model.py:356-360, nnet/layers.py:29-76, 82-503)."""
import ctypes as C
import weakref

import torch

from . import _lib as L
from . import ops

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def dense_strides(shape):
    st, n = [], 1
    for d in reversed(shape):
        st.append(n)
        n *= int(d)
    return tuple(reversed(st))


class _Entry:
    __slots__ = ("key", "params", "ver", "dst", "jobs", "epoch", "key_build")


class WeightPlan:
    def __init__(self):
        self.entries = {}
        self.order = []
        self.table = None          # device uint8 tensor holding the avec_copy_job array of every entry
        self.table_n = 0
        self.table_total = 0
        self.table_dirty = True
        self.epoch = 0             # bumped by invalidate(): every entry is stale

    def invalidate(self):
        self.epoch += 1

    def stale(self):
        return bool(self.order) and any(self.entries[k].epoch != self.epoch for k in self.order[:1] + self.order[-1:])

    def clear(self):
        self.__init__()

    @staticmethod
    def _ver(params):
        return tuple((p._version, p.data_ptr()) for p in params)

    @staticmethod
    def _alive(e):
        ps = [r() for r in e.params]
        return None if any(p is None for p in ps) else ps

    def _purge(self):
        """forget layouts whose parameters were freed (entries hold weak references only)"""
        dead = [k for k in self.order if self._alive(self.entries[k]) is None]
        if dead:
            for k in dead:
                del self.entries[k]
            self.order = [k for k in self.order if k in self.entries]
            self.table_dirty = True

    def _make_jobs(self, e, segs, cols, ld):
        def pitched(s):
            assert s % cols == 0 or s < cols, "destination stride must stay inside a row or step whole rows"
            return (s // cols) * ld + (s % cols)
        jobs = []
        for view, dst_off, dst_strides in segs:
            assert view.dim() <= 4 and view.dtype in _DT
            n = [1] * (4 - view.dim()) + [int(x) for x in view.shape]
            ss = [0] * (4 - view.dim()) + [int(x) for x in view.stride()]
            ds = [0] * (4 - view.dim()) + [pitched(int(x)) for x in dst_strides]
            jobs.append((view, pitched(int(dst_off)), n, ss, ds))
        return jobs

    def _upload(self, entries):
        """host job table -> device tensor (not capturable: happens on first use / after a parameter moved, i.e. during warm-up)"""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("avec_b200.weights: a new weight layout was requested during CUDA-graph capture; run one eager step first")
        njobs = sum(len(e.jobs) for e in entries)
        arr = (L.CopyJob * njobs)()
        k, start, chunks = 0, 0, []
        for e in entries:
            esz = e.dst.element_size()
            for view, dst_off, n, ss, ds in e.jobs:
                j = arr[k]
                j.src, j.dst, j.start = view.data_ptr(), e.dst.data_ptr() + dst_off * esz, start
                for q in range(4):
                    j.n[q], j.ss[q], j.ds[q] = n[q], ss[q], ds[q]
                j.src_dtype, j.dst_dtype = _DT[view.dtype], _DT[e.dst.dtype]
                numel = n[0] * n[1] * n[2] * n[3]
                assert numel < 2 ** 31
                chunks += [(k, c) for c in range(-(-numel // L.COPY_CHUNK))]
                start += numel
                k += 1
        dev = entries[0].dst.device
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        chunk_t = torch.tensor(chunks, dtype=torch.int32).to(dev)
        return (table, chunk_t), njobs, len(chunks)

    def _run(self, table, njobs, nchunks):
        L.check(L.load().avec_convert_multi(table[0].data_ptr(), table[1].data_ptr(), nchunks, ops._stream()), "avec_convert_multi")

    def refresh_all(self):
        self._purge()
        if not self.order:
            return
        entries = [self.entries[k] for k in self.order]
        moved = False
        for e in entries:
            v = self._ver(self._alive(e))
            if tuple(x[1] for x in v) != tuple(x[1] for x in e.ver):
                moved = True                     # a parameter's storage moved (flat optimizer buffers): source pointers changed
                self._rebind(e)
            e.ver = v
        if self.table_dirty or moved:
            self.table, self.table_n, self.table_total = self._upload(entries)
            self.table_dirty = False
        self._run(self.table, self.table_n, self.table_total)
        for e in entries:
            e.epoch = self.epoch

    def _rebind(self, e):
        """re-derive the source views of an entry after its parameters' storage moved"""
        rows, cols, segs, pad = e.key_build(*[p.detach() for p in self._alive(e)])
        e.jobs = self._make_jobs(e, segs, cols, e.dst.stride(0) if e.dst.dim() == 2 else cols)

    def layout(self, params, tag, build, dtype):
        """build(*detached params) -> (rows, cols, segments, pad): segments = [(view of a parameter (<= 4-d), destination element
        offset, destination strides per view dim)] in terms of a DENSE [rows, cols] destination; pad: pitch the rows for TMA."""
        key = (tuple(id(p) for p in params), tag, dtype)
        e = self.entries.get(key)
        if e is not None:
            ps = self._alive(e)
            if ps is None or any(a is not b for a, b in zip(ps, params)):      # id() reused by a new parameter
                del self.entries[key]
                self.order.remove(key)
                self.table_dirty = True
                e = None
        if e is None:
            rows, cols, segs, pad = build(*[p.detach() for p in params])
            dev = params[0].device
            if dev.type != "cuda":
                raise RuntimeError("avec_b200 weight layouts need CUDA parameters: the hot path has no CPU fallback")
            ld = ops.row_pitch(cols, dtype) if pad else cols
            full = torch.zeros((rows, ld), device=dev, dtype=dtype)
            e = _Entry()
            e.key, e.params, e.dst = key, [weakref.ref(p) for p in params], (full[:, :cols] if ld != cols else full)
            e.key_build = build
            e.jobs = self._make_jobs(e, segs, cols, ld)
            e.ver, e.epoch = self._ver(params), self.epoch
            self.entries[key] = e
            self.order.append(key)
            self.table_dirty = True
            t, n, tot = self._upload([e])
            self._run(t, n, tot)
            return e.dst
        if e.epoch != self.epoch:
            self.refresh_all()
        v = self._ver(params)
        if v != e.ver:
            if tuple(x[1] for x in v) != tuple(x[1] for x in e.ver):
                self._rebind(e)
                self.table_dirty = True
            e.ver = v
            t, n, tot = self._upload([e])
            self._run(t, n, tot)
        return e.dst


PLAN = WeightPlan()


# ------------------------------------------------------------------------------------------------ layout builders
def b_plain(fn=None, pad=True):
    """2-d [rows, cols] copy of fn(param) (a VIEW of the parameter, <= 4-d; default: the parameter reshaped to [N, -1])"""
    def build(p):
        v = fn(p) if fn is not None else p.reshape(p.shape[0], -1)
        assert v.untyped_storage().data_ptr() == p.untyped_storage().data_ptr(), "layout functions must return views of the parameter"
        rows = int(v.shape[0])
        cols = 1
        for x in v.shape[1:]:
            cols *= int(x)
        return rows, cols, [(v, 0, dense_strides(v.shape))], pad
    return build


def b_cat(pad=True):
    """parameters stacked along rows, each reshaped to [N_i, K]"""
    def build(*ps):
        vs = [p.reshape(p.shape[0], -1) for p in ps]
        cols = int(vs[0].shape[1])
        segs, off = [], 0
        for v in vs:
            segs.append((v, off, (cols, 1)))
            off += int(v.shape[0]) * cols
        return off // cols, cols, segs, pad
    return build


def b_heads_rows(H, d, dp, pad=True):
    """[H*d, K] (or [H*d]) parameters stacked along rows with every head zero-padded to dp rows: [n*H*dp, K]"""
    def build(*ps):
        K = int(ps[0].reshape(ps[0].shape[0], -1).shape[1])
        segs, off = [], 0
        for p in ps:
            segs.append((p.reshape(H, d, K), off, (dp * K, K, 1)))
            off += H * dp * K
        return len(ps) * H * dp, K, segs, pad
    return build


def b_heads_cols(H, d, dp, pad=True):
    """[N, H*d] parameter with every head zero-padded to dp columns: [N, H*dp]"""
    def build(p):
        N = int(p.shape[0])
        return N, H * dp, [(p.reshape(N, H, d), 0, (H * dp, dp, 1))], pad
    return build


def b_custom(view_fn, rows, cols, dst_strides, pad=True):
    """fn(param) view copied with explicit destination strides (zero elsewhere), e.g. the visual stem's k = (kt*7+kh)*8 + kw packing"""
    def build(p):
        return rows, cols, [(view_fn(p), 0, dst_strides)], pad
    return build
