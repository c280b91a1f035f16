"""Device-timed stopwatch (CUDA events on the current stream) with a CPU fallback.

reference: src/caffe/util/benchmark.cpp:8-95 (caffe::Timer: cudaEvent pair in GPU mode,
posix_time in CPU mode).
"""
from __future__ import annotations

import time

import torch


class Timer:
    def __init__(self, device=None):
        self.cuda = device is not None and torch.device(device).type == "cuda" and torch.cuda.is_available()
        self.running = False
        self.has_run = False
        self._ms = 0.0
        if self.cuda:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)

    def start(self):
        if not self.running:
            if self.cuda:
                self.e0.record()
            else:
                self.t0 = time.perf_counter()
            self.running = True
            self.has_run = True

    def stop(self):
        if self.running:
            if self.cuda:
                self.e1.record()
                self.e1.synchronize()
                self._ms = self.e0.elapsed_time(self.e1)
            else:
                self._ms = (time.perf_counter() - self.t0) * 1e3
            self.running = False

    def milliseconds(self) -> float:
        if not self.has_run:
            return 0.0
        if self.running:
            self.stop()
        return self._ms

    def seconds(self) -> float:
        return self.milliseconds() / 1e3
