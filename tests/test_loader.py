"""CPU tier: ring logic of medt_b200.loader.PrefetchLoader (device agnostic: pageable buffers, no-op events without CUDA)."""
import threading
import time

import pytest
import torch

from medt_b200.loader import PrefetchLoader


def _batches(n, N=2, S=8):
    return [(torch.full((N, 3, S, S), float(i)), torch.full((N, S, S), i, dtype=torch.long), [f"{i}.png"]) for i in range(n)]


def test_prefetch_loader_yields_every_batch_in_order_and_reuses_slots():
    data = _batches(11)
    feed = PrefetchLoader(data, (2, 3, 8, 8), depth=3, device="cpu")
    seen, ptrs = [], set()
    for x, y, rest in feed:
        seen.append((float(x[0, 0, 0, 0]), int(y[0, 0, 0]), rest[0][0]))
        ptrs.add(x.data_ptr())
        time.sleep(0.002)                      # let the producer run ahead: it must block on the ring, not overwrite
        assert float(x[1, 2, 7, 7]) == seen[-1][0]
    assert seen == [(float(i), i, f"{i}.png") for i in range(11)]
    assert len(ptrs) == 3                      # the ring's three staging buffers, recycled
    assert len(list(feed)) == 11               # re-iterable (one pass per epoch)


def test_prefetch_loader_drops_or_passes_ragged_batches_and_propagates_errors():
    data = _batches(3) + [(torch.zeros(1, 3, 8, 8), torch.zeros(1, 8, 8, dtype=torch.long), ["tail.png"])]
    feed = PrefetchLoader(data, (2, 3, 8, 8), depth=2, device="cpu")
    assert len(list(feed)) == 3 and feed.batches_dropped == 1
    keep = PrefetchLoader(data, (2, 3, 8, 8), depth=2, device="cpu", drop_last=False)
    shapes = [tuple(x.shape) for x, _, _ in keep]
    assert shapes[-1] == (1, 3, 8, 8) and len(shapes) == 4

    def broken():
        yield data[0]
        raise OSError("disk gone")
    with pytest.raises(OSError, match="disk gone"):
        list(PrefetchLoader(broken(), (2, 3, 8, 8), depth=2, device="cpu"))
    with pytest.raises(ValueError):
        PrefetchLoader(data, (2, 3, 8, 8), depth=1, device="cpu")


def test_prefetch_loader_early_exit_stops_the_producer():
    n0 = threading.active_count()
    feed = PrefetchLoader(_batches(50), (2, 3, 8, 8), depth=2, device="cpu")
    for i, _ in enumerate(feed):
        if i == 3:
            break
    time.sleep(0.2)
    assert threading.active_count() <= n0 + 1
