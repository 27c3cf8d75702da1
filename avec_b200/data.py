"""Host -> device input pipeline of the training step ("async pinned H2D").

The reference moves every batch with a blocking, pageable copy (functions.py:118 pin_memory=False, nnet/model.py:737-745).
PinnedPrefetcher keeps the NEXT batch's copy in flight while the current step computes: batches are staged in pinned host memory,
copied on a dedicated copy stream into one of two device buffer sets, and handed to the compute stream with an event wait - the
step never waits for PCIe unless the copy is slower than the step itself."""
import torch


class PinnedPrefetcher:
    def __init__(self, batches, device, depth=2):
        """batches: iterable of dicts name -> CPU tensor (same shapes every step).  Iterating yields dicts of device tensors that
        stay valid until the next-but-one `next()` (double buffering)."""
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.slots, self.pinned, self.events = [None] * depth, [None] * depth, [None] * depth
        self.head = 0
        self.inflight = []
        for _ in range(depth - 1):
            self._issue()

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            return False
        s = self.head % self.depth
        self.head += 1
        if self.slots[s] is None:
            self.slots[s] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in batch.items()}
            self.pinned[s] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in batch.items()}
        # the compute stream may still be reading this slot from two steps ago
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        for k, v in batch.items():
            src = v if v.is_pinned() else self.pinned[s][k].copy_(v)
            with torch.cuda.stream(self.copy_stream):
                self.slots[s][k].copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        self.inflight.append((s, ev))
        return True

    def __iter__(self):
        return self

    def __next__(self):
        self._issue()
        if not self.inflight:
            raise StopIteration
        s, ev = self.inflight.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        return self.slots[s]

    def bytes_per_step(self):
        s = next((x for x in self.slots if x is not None), None)
        return sum(v.numel() * v.element_size() for v in s.values()) if s else 0
