"""Out-of-core sample stream: an IO thread walks a list of files (for several passes), cuts them into blocks at line
boundaries and keeps ``num_buffers`` parsed blocks ahead of the consumer (reference: ps/src/ml/disk_stream/
{disk_streamer,disk_reader,multi_buffer}.hpp — IO thread + rotating buffers + LibSVM parser).

``get_next_data(n)`` returns up to ``n`` samples as ``(SparseBatch, labels)``; an empty batch signals the end."""
from __future__ import annotations

import glob
import os
import queue
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .features import SparseBatch


class DiskStreamer:
    def __init__(self, files: Sequence[str] | str, feature_dim: int, num_passes: int = 1, num_buffers: int = 2,
                 block_bytes: int = 4 << 20, feature_one_based: bool = False, label_one_based: bool = False,
                 snappy_compressed: bool = False, parser_threads: int = 0):
        if isinstance(files, str):
            files = sorted(glob.glob(os.path.join(files, "*"))) if os.path.isdir(files) else [files]
        if not files:
            raise ValueError("DiskStreamer: no input files")
        self.files, self.feature_dim, self.num_passes = list(files), feature_dim, num_passes
        self.block_bytes = block_bytes
        self.opts = (feature_one_based, label_one_based)
        self.snappy = snappy_compressed
        self.parser_threads = parser_threads
        self.q: "queue.Queue" = queue.Queue(maxsize=max(1, num_buffers))
        self.stop_flag = threading.Event()
        self.pending: Optional[Tuple[np.ndarray, ...]] = None
        self.done = False
        self.lock = threading.Lock()
        self.thread = threading.Thread(target=self._io_loop, name="disk-streamer", daemon=True)
        self.thread.start()

    # ---- IO thread ---------------------------------------------------------------------------------------------
    def _blocks(self, path: str):
        if self.snappy:                                       # compressed files are one snappy block each
            from ..data import native
            with open(path, "rb") as f:
                yield native.module().snappy_uncompress(f.read())
            return
        carry = b""
        with open(path, "rb") as f:
            while True:
                chunk = f.read(self.block_bytes)
                if not chunk:
                    break
                buf = carry + chunk
                cut = buf.rfind(b"\n")
                if cut < 0:
                    carry = buf
                    continue
                carry = buf[cut + 1:]
                yield buf[: cut + 1]
        if carry.strip():
            yield carry

    def _io_loop(self):
        from ..data import native
        parse = native.module().parse_libsvm
        try:
            for _ in range(self.num_passes):
                for path in self.files:
                    for block in self._blocks(path):
                        if self.stop_flag.is_set():
                            return
                        item = parse(block, self.opts[0], self.opts[1], -1, self.parser_threads)
                        while not self.stop_flag.is_set():
                            try:
                                self.q.put(item, timeout=0.1)
                                break
                            except queue.Full:
                                continue
            self._put_final(None)
        except Exception as e:                                # surfaced to the consumer
            self._put_final(e)

    def _put_final(self, item):
        while not self.stop_flag.is_set():
            try:
                self.q.put(item, timeout=0.1)
                return
            except queue.Full:
                continue

    # ---- consumer ----------------------------------------------------------------------------------------------
    def get_next_data(self, num_data: int) -> Tuple[SparseBatch, torch.Tensor]:
        """Thread-safe.  Fewer than ``num_data`` samples are returned only at the end of the stream."""
        labs: List[np.ndarray] = []
        ptrs: List[np.ndarray] = []
        idxs: List[np.ndarray] = []
        vals: List[np.ndarray] = []
        need = num_data
        with self.lock:
            while need > 0 and not self.done:
                if self.pending is None:
                    item = self.q.get()
                    if item is None:
                        self.done = True
                        break
                    if isinstance(item, Exception):
                        self.done = True
                        raise item
                    self.pending = item
                lab, ptr, idx, val = self.pending
                take = min(need, lab.shape[0])
                a, b = int(ptr[0]), int(ptr[take])
                labs.append(lab[:take])
                ptrs.append(ptr[1: take + 1] - ptr[0])
                idxs.append(idx[a:b])
                vals.append(val[a:b])
                need -= take
                self.pending = None if take == lab.shape[0] else (lab[take:], ptr[take:] - ptr[take], idx[b:], val[b:])
        indptr = [np.zeros(1, np.int64)]
        base = 0
        for p in ptrs:
            indptr.append(p + base)
            base += int(p[-1]) if p.size else 0
        cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt))
        batch = SparseBatch(np.concatenate(indptr), cat(idxs, np.int32).astype(np.int64), cat(vals, np.float32),
                            self.feature_dim)
        return batch, torch.from_numpy(cat(labs, np.int32).copy())

    def shutdown(self):
        self.stop_flag.set()
        try:
            while True:
                self.q.get_nowait()
        except queue.Empty:
            pass
        self.thread.join(timeout=5)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.shutdown()
