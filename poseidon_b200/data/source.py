"""Batch sources + the pinned-memory / copy-stream prefetcher behind the data layers.

A *source* yields host batches ``(data[N,C,H,W] uint8|float32, label[N] float32)``;
the :class:`Prefetcher` runs it on a background thread, stages batches in a ring of
pinned buffers and issues the H2D copy on a dedicated CUDA stream so that the copy of
step i+1 overlaps the compute of step i (the reference's single prefetch thread + sync
``cudaMemcpy``: src/caffe/layers/base_data_layer.cpp:57-104, base_data_layer.cu:9-22).
"""
from __future__ import annotations

import queue
import threading
from typing import Optional, Tuple

import numpy as np
import torch



class SyntheticSource:
    """Endless synthetic uint8 images from a small pre-generated pool (ImageNet-shaped
    benchmarking without a dataset; the reference's hook for this is DummyData)."""

    def __init__(self, batch, shape, num_classes=1000, pool=4, seed=0, dtype=torch.uint8):
        g = torch.Generator().manual_seed(seed)
        c, h, w = shape
        self.batch = batch
        self.pool = []
        for _ in range(pool):
            if dtype == torch.uint8:
                x = torch.randint(0, 256, (batch, c, h, w), generator=g, dtype=torch.uint8)
            else:
                x = torch.rand((batch, c, h, w), generator=g, dtype=torch.float32)
            y = torch.randint(0, num_classes, (batch,), generator=g).float()
            if torch.cuda.is_available():
                x, y = x.pin_memory(), y.pin_memory()      # batches are handed to the copy engine as-is
            self.pool.append((x, y))
        self.i = 0

    def next_batch(self):
        b = self.pool[self.i % len(self.pool)]
        self.i += 1
        return b


class DBSource:
    """Sequential reader over a record DB with worker sharding and wrap-around.
    reference: src/caffe/layers/data_layer.cpp:143-259."""

    def __init__(self, reader, batch, offset=0, stride=1, rand_skip=0, seed=None):
        self.reader, self.batch = reader, batch
        self.stride = max(1, stride)
        n = len(reader)
        if n == 0:
            raise ValueError("empty database")
        self.pos = offset % n
        if rand_skip:
            rng = np.random.RandomState(seed)
            self.pos = (self.pos + int(rng.randint(0, rand_skip)) * self.stride) % n
        d = reader.datum(self.pos)
        self.shape = (d.channels, d.height, d.width)
        self.is_bytes = d.has("data") and len(d.data) > 0

    def next_batch(self):
        c, h, w = self.shape
        n = len(self.reader)
        if self.is_bytes:
            out = np.empty((self.batch, c, h, w), dtype=np.uint8)
        else:
            out = np.empty((self.batch, c, h, w), dtype=np.float32)
        lab = np.zeros((self.batch,), dtype=np.float32)
        for i in range(self.batch):
            d = self.reader.datum(self.pos)
            if self.is_bytes:
                out[i] = np.frombuffer(d.data, dtype=np.uint8).reshape(c, h, w)
            else:
                out[i] = np.asarray(d.float_data, dtype=np.float32).reshape(c, h, w)
            lab[i] = d.label or 0
            self.pos = (self.pos + self.stride) % n
        return torch.from_numpy(out), torch.from_numpy(lab)


class ArraySource:
    """Batches from in-memory arrays (MemoryData / .npz files)."""

    def __init__(self, data, labels, batch, offset=0, stride=1):
        self.data = torch.as_tensor(data)
        self.labels = torch.as_tensor(labels).float().reshape(len(self.data), -1)
        if len(self.data) % 1:
            raise ValueError
        self.batch = batch
        idx = torch.arange(offset, len(self.data), stride)
        self.idx = idx
        self.pos = 0

    def next_batch(self):
        n = len(self.idx)
        sel = self.idx[(torch.arange(self.batch) + self.pos) % n]
        self.pos = (self.pos + self.batch) % n
        lab = self.labels[sel]
        return self.data[sel], lab.squeeze(1) if lab.shape[1] == 1 else lab


class Prefetcher:
    """Background producer: source → pinned ring → (copy stream) → device."""

    def __init__(self, source, device, depth: int = 2):
        self.source = source
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth = depth
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.stop = False
        self.err: Optional[BaseException] = None
        self.h2d_bytes = 0
        if self.cuda:
            self.stream = torch.cuda.Stream(device=self.device)
            self.ring = []            # (pinned_x, pinned_y, event)
            self.slot = 0
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _stage(self, x, y):
        if not self.cuda:
            return x, y, None
        if x.is_pinned() and y.is_pinned():
            # source already produced page-locked batches (its own ring): copy straight from them
            ev = torch.cuda.Event()
            with torch.cuda.stream(self.stream):
                gx = x.to(self.device, non_blocking=True)
                gy = y.to(self.device, non_blocking=True)
                ev.record(self.stream)
            # the source recycles its page-locked buffers a couple of batches later: make sure the previous
            # batch's copy has drained before asking for more (it almost always has — costs nothing)
            prev, self._prev_pinned_ev = getattr(self, "_prev_pinned_ev", None), ev
            if prev is not None:
                prev.synchronize()
            return gx, gy, ev
        n_slots = self.depth + 2
        if len(self.ring) < n_slots:
            px = torch.empty(x.shape, dtype=x.dtype).pin_memory()
            py = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            ev = torch.cuda.Event()
            self.ring.append([px, py, ev, False])
        slot = self.ring[self.slot % n_slots]
        self.slot += 1
        px, py, ev, used = slot
        if used:
            ev.synchronize()          # the copy that last read this pinned buffer is done
        if px.shape != x.shape or px.dtype != x.dtype:
            px = slot[0] = torch.empty(x.shape, dtype=x.dtype).pin_memory()
            py = slot[1] = torch.empty(y.shape, dtype=y.dtype).pin_memory()
        px.copy_(x)
        py.copy_(y)
        with torch.cuda.stream(self.stream):
            gx = px.to(self.device, non_blocking=True)
            gy = py.to(self.device, non_blocking=True)
            ev.record(self.stream)
        slot[3] = True
        return gx, gy, ev

    def _run(self):
        try:
            if self.cuda:
                torch.cuda.set_device(self.device)
            while not self.stop:
                x, y = self.source.next_batch()
                item = self._stage(x, y)
                while not self.stop:
                    try:
                        self.q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:  # surfaced on the consumer side
            self.err = e
            self.q.put(None)

    def next(self) -> Tuple[torch.Tensor, torch.Tensor]:
        item = self.q.get()
        if item is None:
            raise RuntimeError("data prefetch thread failed") from self.err
        x, y, ev = item
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            x.record_stream(torch.cuda.current_stream(self.device))
            y.record_stream(torch.cuda.current_stream(self.device))
        self.h2d_bytes = x.numel() * x.element_size() + y.numel() * y.element_size()
        return x, y

    def close(self):
        self.stop = True
        try:
            while True:
                self.q.get_nowait()
        except queue.Empty:
            pass
        self.thread.join(timeout=2)
