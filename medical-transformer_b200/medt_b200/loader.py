"""Host side of the training loop: the reference feeds the GPU from `DataLoader(dataset, batch_size,
shuffle=True)` with num_workers=0 and pageable tensors (train.py:88-91, :137-138: a blocking `X_batch.to('cuda')` per
step).  `PrefetchLoader` wraps any iterable of `(x, y, *rest)` host batches (that DataLoader, unchanged) and keeps the GPU
fed without host synchronisation:

  * a background thread pulls the next batches from the wrapped iterable (decode / augmentation overlap the GPU step) and
    writes them into a RING of pinned staging buffers;
  * each ring slot carries a CUDA event recorded after the slot's host->device copy was enqueued; the producer waits for
    that event before it overwrites the slot, so a DMA never reads a half-rewritten buffer (the race ADVICE r1 found in
    TrainStep's single staging buffer cannot occur here by construction);
  * iteration yields `(x_pinned, y_pinned, rest)`; `TrainStep.step(x_pinned, y_pinned)` takes its pinned fast path (async
    H2D straight from the ring); the next `next()` records the slot's event on the current stream and frees the slot.

Ragged last batches are dropped (the captured step has a fixed shape; `drop_last=False` yields them un-pinned instead).
The ring logic is device agnostic: without CUDA the buffers are pageable and the events are no-ops (CPU tests).
"""
from __future__ import annotations

import queue
import threading

import torch


class _NullEvent:
    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class PrefetchLoader:
    def __init__(self, loader, batch_shape, depth: int = 3, device=None, drop_last: bool = True, y_dtype=torch.long):
        self.loader, self.depth, self.drop_last = loader, int(depth), drop_last
        if self.depth < 2:
            raise ValueError("PrefetchLoader needs a ring of at least 2 slots")
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        self.device = torch.device(device) if device is not None else (torch.device("cuda") if self.cuda else torch.device("cpu"))
        n, c, h, w = batch_shape
        mk = (lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype).pin_memory()) if self.cuda else \
             (lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype))
        self.xs = [mk(n, c, h, w) for _ in range(self.depth)]
        self.ys = [mk(n, h, w, dtype=y_dtype) for _ in range(self.depth)]
        self.events = [torch.cuda.Event() if self.cuda else _NullEvent() for _ in range(self.depth)]
        self.batches_dropped = 0

    def __len__(self):
        return len(self.loader)

    def _producer(self, q, free, stop):
        try:
            for batch in self.loader:
                if stop.is_set():
                    break
                x, y, *rest = batch
                if tuple(x.shape) != tuple(self.xs[0].shape) or tuple(y.shape) != tuple(self.ys[0].shape):
                    if self.drop_last:
                        self.batches_dropped += 1
                        continue
                    q.put(("ragged", x, y, rest))
                    continue
                slot = free.get()
                if slot is None:
                    break
                self.events[slot].synchronize()            # the previous DMA out of this slot has completed
                self.xs[slot].copy_(x)
                self.ys[slot].copy_(y)
                q.put(("slot", slot, None, rest))
            q.put(("end", None, None, None))
        except BaseException as e:  # noqa: BLE001 -- surfaced in the consumer thread
            q.put(("error", e, None, None))

    def __iter__(self):
        q: queue.Queue = queue.Queue()
        free: queue.Queue = queue.Queue()
        for s in range(self.depth):
            free.put(s)
        stop = threading.Event()
        th = threading.Thread(target=self._producer, args=(q, free, stop), daemon=True)
        th.start()
        pending = None

        def release(slot):
            if self.cuda:
                self.events[slot].record(torch.cuda.current_stream(self.device))
            free.put(slot)

        try:
            while True:
                if pending is not None:                     # the consumer has enqueued its copies of the previous slot by now
                    release(pending)
                    pending = None
                kind, a, b, rest = q.get()
                if kind == "end":
                    return
                if kind == "error":
                    raise a
                if kind == "ragged":
                    yield a, b, rest
                    continue
                pending = a
                yield self.xs[a], self.ys[a], rest
        finally:
            stop.set()
            free.put(None)
            if pending is not None:
                release(pending)
