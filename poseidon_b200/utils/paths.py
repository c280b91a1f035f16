"""File-name resolution for prototxt fields (net:, source:, mean_file:, snapshot_prefix:, infogain source:).

The reference's shipped solvers and nets spell paths as ``CAFFE_ROOT/examples/mnist/...`` (``POSEIDON_ROOT/models/...``
under models/) and ask the user to replace the placeholder by hand with the full application directory before launching
(examples/mnist/run_local.py:19-20, examples/mnist/lenet_solver.prototxt:2,
examples/cifar10/cifar10_quick_train_test.prototxt).  Here the placeholder is resolved at load time: ``$CAFFE_ROOT`` /
``$POSEIDON_ROOT`` if set, otherwise the nearest ancestor of the referring file under which the rest of the path exists.
Relative paths are tried against the referring file's directory, its ancestors and the working directory, which covers
both conventions found in the wild (relative to the model directory / relative to the repo root the tools were started
from).
"""
from __future__ import annotations

import os
from typing import Optional

PLACEHOLDERS = ("CAFFE_ROOT", "POSEIDON_ROOT")      # examples/ use the first, models/bvlc_* the second
_root_hint: Optional[str] = None      # root discovered by an earlier successful expansion in this process


def _ancestors(d: Optional[str], depth: int = 5):
    out = []
    d = os.path.abspath(d) if d else None
    while d and len(out) < depth:
        out.append(d)
        parent = os.path.dirname(d)
        if parent == d:
            break
        d = parent
    return out


def split_placeholder(path: str):
    """("CAFFE_ROOT", "x/y") for "CAFFE_ROOT/x/y"; (None, path) otherwise."""
    for ph in PLACEHOLDERS:
        if path == ph or (path or "").startswith(ph + "/"):
            return ph, path[len(ph):].lstrip("/")
    return None, path


def expand_placeholder(path: str, model_dir: Optional[str] = None, must_exist: bool = True) -> str:
    """``CAFFE_ROOT/x/y`` -> ``<root>/x/y``.  Paths without a placeholder are returned unchanged."""
    global _root_hint
    ph, rest = split_placeholder(path)
    if ph is None:
        return path
    env = os.environ.get(ph)
    if env:
        return os.path.join(env, rest)
    roots = ([_root_hint] if _root_hint else []) + _ancestors(model_dir) + [os.getcwd()]
    for r in roots:                                        # the target itself
        if os.path.exists(os.path.join(r, rest)):
            _root_hint = _root_hint or r
            return os.path.join(r, rest)
    if not must_exist:
        for r in roots:                                    # a file about to be written: its directory
            if os.path.isdir(os.path.dirname(os.path.join(r, rest))):
                return os.path.join(r, rest)
    return os.path.join(roots[0], rest) if roots else rest


def resolve(path: str, model_dir: Optional[str] = None, must_exist: bool = True, basename_fallback: bool = False) -> str:
    """Best existing candidate for a path named in a prototxt; the (expanded) input if nothing exists."""
    if not path:
        return path
    ph, rest = split_placeholder(path)
    if ph is not None:
        full = expand_placeholder(path, model_dir, must_exist)
        if os.path.exists(full) or not must_exist:
            return full
        fallback, path = full, rest                        # keep looking for the remainder as a relative path
    else:
        fallback = path
    if os.path.isabs(path) or os.path.exists(path) or not model_dir:
        return path if (os.path.isabs(path) or os.path.exists(path)) else fallback
    bases = _ancestors(model_dir, 3) + [os.getcwd()]
    for base in bases:
        cand = os.path.join(base, path)
        if os.path.exists(cand):
            return cand
    if basename_fallback:                                   # a solver next to its net, named with a stale directory
        for base in bases:
            cand = os.path.join(base, os.path.basename(path))
            if os.path.exists(cand):
                return cand
    return fallback
