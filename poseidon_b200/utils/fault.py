"""Fault-injection hooks for the distributed tests (the reference is fail-fast with no injection).

``POSEIDON_FAULT`` holds ``;``-separated directives, each ``kind:key=value,...``:

    delay:rank=1,step=3,ms=200        sleep 200 ms on rank 1 before iteration 3 (straggler; exercises SSP staleness
                                      and the bounded device-side flag waits)
    kill:rank=2,step=5                hard-exit rank 2 before iteration 5 (peers must abort through the watchdog /
                                      collective timeout instead of hanging forever)
    raise:rank=0,step=2               raise RuntimeError on rank 0 (exercises the snapshot-and-restart path)

A directive with ``attempt=k`` fires only in the k-th start of a supervised job (``tools.launch --max_restarts`` exports
``POSEIDON_ATTEMPT``), so a restarted job does not die again at the same step.

``maybe_inject(rank, step)`` is called by ``Solver`` at the top of every training iteration; with the variable unset
it is a dictionary lookup.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Tuple

_PARSED: List[Tuple[str, Dict[str, int]]] | None = None


def parse(spec: str) -> List[Tuple[str, Dict[str, int]]]:
    out = []
    for item in filter(None, (s.strip() for s in spec.split(";"))):
        kind, _, rest = item.partition(":")
        kv = {}
        for pair in filter(None, rest.split(",")):
            k, _, v = pair.partition("=")
            kv[k.strip()] = int(v)
        if kind not in ("delay", "kill", "raise"):
            raise ValueError(f"POSEIDON_FAULT: unknown directive '{kind}'")
        out.append((kind, kv))
    return out


def directives():
    global _PARSED
    if _PARSED is None:
        _PARSED = parse(os.environ.get("POSEIDON_FAULT", ""))
    return _PARSED


def reset():
    global _PARSED
    _PARSED = None


def maybe_inject(rank: int, step: int) -> None:
    for kind, kv in directives():
        if kv.get("rank", 0) != rank or kv.get("step", -1) != step:
            continue
        if "attempt" in kv and kv["attempt"] != int(os.environ.get("POSEIDON_ATTEMPT", "0")):
            continue
        if kind == "delay":
            time.sleep(kv.get("ms", 100) / 1e3)
        elif kind == "kill":
            os._exit(kv.get("code", 17))
        elif kind == "raise":
            raise RuntimeError(f"injected fault on rank {rank} at step {step}")
