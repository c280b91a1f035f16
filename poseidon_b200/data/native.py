"""Native (C++) host runtime bindings: record-file batch loader + wire compression.

``csrc_host/record_loader.cpp`` is plain C++17 + pybind11 (no CUDA), compiled in-tree to
``poseidon_b200/_ext/poseidon_b200_host.so`` by :func:`build` (also called from ``__graft_entry__.build``).
:class:`NativeDBSource` is the drop-in replacement of :class:`poseidon_b200.data.source.DBSource`: same cursor
semantics (offset / stride sharding, wrap-around; reference: src/caffe/layers/data_layer.cpp:143-259) but the
records are parsed and copied into page-locked batch buffers by a C++ thread pool running ahead of the trainer.
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys
import sysconfig
import threading

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC_DIR = os.path.join(_ROOT, "csrc_host")
SRCS = [os.path.join(SRC_DIR, "record_loader.cpp"), os.path.join(SRC_DIR, "leveldb_reader.cpp"),
        os.path.join(SRC_DIR, "libsvm_parser.cpp")]
SRC = SRCS[0]
EXT_DIR = os.path.join(_ROOT, "poseidon_b200", "_ext")
# POSEIDON_HOST_SO: use a module built elsewhere (scripts/sanitize_host.sh points it at an ASan / TSan build)
SO = os.environ.get("POSEIDON_HOST_SO") or os.path.join(EXT_DIR, "poseidon_b200_host.so")
_mod = None
_lock = threading.Lock()


def build(force: bool = False, verbose: bool = False) -> str:
    """g++ -O3 -shared the host module (a few seconds); no-op when up to date."""
    os.makedirs(EXT_DIR, exist_ok=True)
    newest = max(os.path.getmtime(f) for f in SRCS + [os.path.join(SRC_DIR, "leveldb_reader.h")])
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    import pybind11
    cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-pthread",
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], "-I", SRC_DIR] + SRCS + ["-o", SO + ".tmp"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    os.replace(SO + ".tmp", SO)
    return SO


def module(build_if_missing: bool = True):
    global _mod
    with _lock:
        if _mod is not None:
            return _mod
        if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in SRCS):
            if not build_if_missing:
                raise RuntimeError(f"{SO} not built")
            build()
        spec = importlib.util.spec_from_file_location("poseidon_b200_host", SO)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["poseidon_b200_host"] = mod
        _mod = mod
        return mod


def available() -> bool:
    try:
        module()
        return True
    except Exception:
        return False


class NativeDBSource:
    """Batches from a PDB record file, produced by the C++ loader into a ring of (pinned) host buffers.

    ``next_batch()`` returns tensors that alias a ring slot; the slot is recycled ``depth`` calls later, which is
    after the :class:`~poseidon_b200.data.source.Prefetcher` has issued (and event-synchronised) its H2D copy.
    """

    def __init__(self, path: str, batch: int, offset: int = 0, stride: int = 1, rand_skip: int = 0, seed=None,
                 threads: int = 0, depth: int = 6, pin: bool | None = None):
        m = module()
        if os.path.isfile(path):
            pdb = path
        else:                                       # a directory: PDB record store or an LMDB environment
            pdb = os.path.join(path, "data.pdb")
            if not os.path.isfile(pdb) and os.path.isfile(os.path.join(path, "data.mdb")):
                pdb = os.path.join(path, "data.mdb")
            elif not os.path.isfile(pdb) and os.path.isfile(os.path.join(path, "CURRENT")):
                pdb = path                          # LevelDB directory
        # decode threads per rank: half the cores of this rank's share of the host (torchrun exports LOCAL_WORLD_SIZE)
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        threads = threads or max(2, min(16, (os.cpu_count() or 4) // (2 * local_world)))
        self.loader = m.BatchLoader(pdb, batch, offset, max(1, stride), threads)
        n = self.loader.num_records()
        if rand_skip:
            import numpy as np
            rng = np.random.RandomState(seed)
            self.loader.seek((offset + int(rng.randint(0, rand_skip)) * max(1, stride)) % n)
        self.batch = batch
        self.shape = tuple(self.loader.shape())
        self.is_bytes = bool(self.loader.is_bytes())
        pin = torch.cuda.is_available() if pin is None else pin
        self.pinned = bool(pin)
        dt = torch.uint8 if self.is_bytes else torch.float32
        self.slots = []
        for _ in range(max(2, depth)):
            x = torch.empty((batch,) + self.shape, dtype=dt)
            y = torch.empty((batch,), dtype=torch.float32)
            if pin:
                x, y = x.pin_memory(), y.pin_memory()
            self.slots.append((x, y))
        self.loader.start([(x.data_ptr(), y.data_ptr()) for x, y in self.slots])
        self._held = []               # slots handed out, oldest first
        self._hold = max(1, len(self.slots) - 2)

    def __len__(self):
        return self.loader.num_records()

    def next_batch(self):
        while len(self._held) >= self._hold:
            self.loader.release(self._held.pop(0))
        s = self.loader.acquire()
        self._held.append(s)
        x, y = self.slots[s]
        if not self.pinned:
            # CPU consumers keep references to what they are handed (prefetch queue, label blobs): give them
            # private copies; only the pinned slots of the CUDA path are handed out in place (copied H2D at once)
            return x.clone(), y.clone()
        return x, y

    def close(self):
        self.loader.stop()

    def __del__(self):
        try:
            self.loader.stop()
        except Exception:
            pass


def f32_to_bf16(src: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Round-to-nearest-even fp32 -> bf16 on the host (wire compression for the CPU/gloo SSP path)."""
    assert src.dtype == torch.float32 and src.is_contiguous() and src.device.type == "cpu"
    out = torch.empty(src.shape, dtype=torch.bfloat16) if out is None else out
    module().f32_to_bf16(src.data_ptr(), out.data_ptr(), src.numel())
    return out


def bf16_to_f32(src: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    assert src.dtype == torch.bfloat16 and src.is_contiguous() and src.device.type == "cpu"
    out = torch.empty(src.shape, dtype=torch.float32) if out is None else out
    module().bf16_to_f32(src.data_ptr(), out.data_ptr(), src.numel())
    return out


class NativeRecordDB:
    """Random-access reader over any database the C++ runtime understands (PDB file, LMDB ``data.mdb``, LevelDB
    directory) with the interface of :class:`poseidon_b200.data.db.RecordReader`."""

    def __init__(self, path: str):
        self.path = path
        self._db = module().RecordDB(path)

    def __len__(self):
        return self._db.size()

    def key(self, i: int) -> bytes:
        return self._db.key(i)

    def value(self, i: int) -> bytes:
        return self._db.value(i)

    def datum(self, i: int):
        from .. import proto as P
        return P.Datum.FromString(self.value(i))

    def __iter__(self):
        for i in range(len(self)):
            yield self.key(i), self.value(i)

    def close(self):
        self._db = None
