"""DWBP scheduler + the library-based (baseline) communication backends.

Distributed wait-free backprop in this framework: every learnable layer is a *bucket*; the
moment autograd has produced all gradients of a bucket (post-accumulate-grad hooks — the
equivalent of "layer i's Backward returned" in the reference's hand-rolled backward loop)
the bucket is handed to a backend which reduces it across ranks and applies the optimizer
step on a communication stream while backprop continues through the lower layers.  The
next forward of layer ℓ waits only on ℓ's own bucket event, not on a global barrier.

Backends in this file (the product's fused NVLink kernels live in ``fused.py``):
  * ``LocalBackend``      – world_size 1: optimizer step only.
  * ``TorchDistBackend``  – NCCL (GPU baseline) / gloo (CPU tests): per-bucket all-reduce from
                            the hooks + unfused step.  "The baseline, not the product."
  * ``SSPBackend``        – bounded-staleness async SGD (Bösen SSP semantics): apply own update
                            now, fold in the other workers' summed updates ≤ s clocks later.

reference: src/caffe/solver.cpp:405-451 (ForwardBackward/DWBP), :455-473 (ThreadSyncWithPS),
:534-540 (JoinSyncThreads), :815-892 (SGD ComputeUpdateValue), blob.cpp:208-248.
"""
from __future__ import annotations

import logging
from collections import deque
from typing import Callable, Dict, List, Optional

import math

import torch

from ..utils.trace import nvtx_range
import torch.distributed as dist

from ..ops import reference as R

log = logging.getLogger("poseidon_b200")

SGD, NESTEROV, ADAGRAD = 0, 1, 2


class Hyper:
    """Per-iteration optimizer hyper-parameters (global part)."""

    def __init__(self, solver_type=SGD, momentum=0.0, weight_decay=0.0, l1=False, delta=1e-8):
        self.solver_type = solver_type
        self.momentum = momentum
        self.weight_decay = weight_decay
        self.l1 = l1
        self.delta = delta
        self.lr = 0.0


class Bucket:
    def __init__(self, bid, layer_idx, layer_name, layer):
        self.id, self.layer_idx, self.layer_name, self.layer = bid, layer_idx, layer_name, layer
        self.params: List[torch.nn.Parameter] = []
        self.lr_mult: List[float] = []
        self.decay_mult: List[float] = []
        self.history: List[torch.Tensor] = []
        self.pending = 0
        self.mode = "dense"          # "dense" | "sfb"
        self.event = None            # CUDA event recorded after this bucket's update
        self.sfb_event = None        # ... and after the fused SFB kernel of an IP weight
        self.self_updating = set()   # ids of params stepped inside their layer's backward (no .grad appears)
        self.numel = 0


def build_buckets(net) -> List[Bucket]:
    """One bucket per layer that owns learnable blobs (CONV/IP weight+bias ≙ the reference's
    two PS tables per layer: src/caffe/caffe_engine.cpp:79-128)."""
    by_layer: Dict[int, Bucket] = {}
    buckets: List[Bucket] = []
    for p, li, lr, wd in zip(net.params, net.param_layer_idx, net.params_lr, net.params_weight_decay):
        if not p.requires_grad:
            continue
        b = by_layer.get(li)
        if b is None:
            b = Bucket(len(buckets), li, net.layer_names[li], net.layers[li])
            by_layer[li] = b
            buckets.append(b)
        b.params.append(p)
        b.lr_mult.append(lr)
        b.decay_mult.append(wd)
        b.history.append(torch.zeros_like(p))
        b.numel += p.numel()
    return buckets


def apply_rule(hyper: Hyper, w, g, h, lr_mult, decay_mult, decay_scale=1.0):
    """One optimizer step on a single tensor (torch ops; the oracle for the fused kernels)."""
    lr = hyper.lr * lr_mult
    wd = hyper.weight_decay * decay_mult * decay_scale
    with torch.no_grad():
        if hyper.solver_type == SGD:
            R.sgd_step(w, g, h, lr, hyper.momentum, wd, hyper.l1)
        elif hyper.solver_type == NESTEROV:
            R.nesterov_step(w, g, h, lr, hyper.momentum, wd, hyper.l1)
        else:
            R.adagrad_step(w, g, h, lr, hyper.delta, wd, hyper.l1)


def compute_update(hyper: Hyper, w, g, h, lr_mult, decay_mult, decay_scale=1.0):
    """Return the step Δ (so that w_new = w − Δ) and advance the history, without touching w."""
    lr = hyper.lr * lr_mult
    wd = hyper.weight_decay * decay_mult * decay_scale
    with torch.no_grad():
        if wd:
            g = g + wd * (torch.sign(w) if hyper.l1 else w)
        if hyper.solver_type == SGD:
            h.mul_(hyper.momentum).add_(g, alpha=lr)
            return h.clone()
        if hyper.solver_type == NESTEROV:
            h_old = h.clone()
            h.mul_(hyper.momentum).add_(g, alpha=lr)
            return (1 + hyper.momentum) * h - hyper.momentum * h_old
        h.add_(g * g)
        return lr * g / (h.sqrt() + hyper.delta)


def foreach_sgd_step(hyper: Hyper, entries, lr_t: torch.Tensor, decay_scale=1.0, gscale=1.0):
    """Caffe's SGD rule (h <- mom*h + lr*(g + wd*w); w <- w - h) over many tensors with ``torch._foreach`` kernels and a
    DEVICE-resident learning rate, so the step can sit inside a captured CUDA graph.  This is the optimizer of the
    constructed vendor baseline (cuDNN/cuBLAS through PyTorch + foreach SGD); the product fuses the step into its
    reduce / wgrad kernels instead.  ``entries``: (param, grad, history, lr_mult, decay_mult); grads are consumed."""
    if hyper.solver_type != SGD or hyper.l1:
        raise RuntimeError("the foreach (CUDA-graph) optimizer of the library backends implements SGD with L2 decay")
    groups: Dict[tuple, list] = {}
    for e in entries:
        groups.setdefault((float(e[3]), float(e[4])), []).append(e)
    with torch.no_grad():
        for (lm, dm), es in groups.items():
            ws = [e[0].data for e in es]
            gs = [e[1] if e[1].dtype == e[0].dtype else e[1].to(e[0].dtype) for e in es]
            hs = [e[2] for e in es]
            if gscale != 1.0:
                torch._foreach_mul_(gs, gscale)
            wd = hyper.weight_decay * dm * decay_scale
            if wd:
                torch._foreach_add_(gs, ws, alpha=wd)
            torch._foreach_mul_(gs, lr_t * lm if lm != 1.0 else lr_t)
            torch._foreach_mul_(hs, hyper.momentum)
            torch._foreach_add_(hs, gs)
            torch._foreach_sub_(ws, hs)


class Backend:
    name = "base"
    uses_comm_stream = False
    deferred_step = False     # library backends: reduce per bucket (DWBP), ONE foreach optimizer step at the end of the iteration
    lr_t: Optional[torch.Tensor] = None

    def enable_deferred_step(self, device):
        """Switch to the CUDA-graph-safe optimizer: learning rate on the device, one foreach step per iteration."""
        self.deferred_step = True
        self.lr_t = torch.zeros((), dtype=torch.float32, device=device)
        self._deferred = []

    def set_lr(self, lr: float):
        if self.lr_t is not None:
            self.lr_t.fill_(float(lr))

    def setup(self, sync: "GradSync"):
        self.sync = sync

    def launch(self, bucket: Bucket):
        raise NotImplementedError

    def finish_iteration(self):
        pass

    def bytes_on_wire(self) -> Dict[str, int]:
        return {}


class LocalBackend(Backend):
    name = "local"

    def launch(self, bucket):
        hy = self.sync.hyper
        if self.deferred_step:
            self._deferred += [(p, p.grad, h, lm, dm) for p, h, lm, dm in
                               zip(bucket.params, bucket.history, bucket.lr_mult, bucket.decay_mult)]
            return
        for p, h, lm, dm in zip(bucket.params, bucket.history, bucket.lr_mult, bucket.decay_mult):
            apply_rule(hy, p.data, p.grad, h, lm, dm)

    def finish_iteration(self):
        if self.deferred_step and self._deferred:
            foreach_sgd_step(self.sync.hyper, self._deferred, self.lr_t)
            self._deferred = []


class TorchDistBackend(Backend):
    """Per-bucket ``all_reduce`` (NCCL on GPUs, gloo on CPU) launched from the backward hooks
    on a side stream, followed by an unfused step.  ``reduce='sum'`` reproduces the reference's
    summed-update semantics (effective LR ∝ #workers); weight decay is scaled by
    world_size so the result equals the PS's sum of per-worker updates under BSP."""
    name = "torchdist"

    def __init__(self, reduce="sum"):
        self.reduce = reduce
        self.dense_bytes = 0

    def setup(self, sync):
        super().setup(sync)
        dev = sync.rank_ctx.device
        self.cuda = dev.type == "cuda"
        self.uses_comm_stream = self.cuda
        self.stream = torch.cuda.Stream(device=dev, priority=-1) if self.cuda else None

    def _reduce_and_step(self, bucket):
        hy = self.sync.hyper
        ws = self.sync.rank_ctx.world_size
        # SFB weights already hold the global gradient (reconstructed from factors)
        red = [p for p in bucket.params if not getattr(p, "_grad_is_global", False)]
        grads = {}
        if red:
            flat = torch.cat([p.grad.reshape(-1) for p in red]) if len(red) > 1 else red[0].grad.reshape(-1)
            if getattr(self, "wire_bf16", False):          # DenseFloat16-style wire format: half the bytes, fp32 at rest
                wire = flat.to(torch.bfloat16)
                dist.all_reduce(wire)
                self.dense_bytes += wire.numel() * wire.element_size()
                flat = wire.float()
            else:
                dist.all_reduce(flat)
                self.dense_bytes += flat.numel() * flat.element_size()
            off = 0
            for p in red:
                grads[id(p)] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        if self.deferred_step:
            # DDP-style: the reduction overlaps backward (comm stream), the optimizer runs once at the end of the step
            self._deferred += [(p, grads.get(id(p), p.grad), h, lm, dm) for p, h, lm, dm in
                               zip(bucket.params, bucket.history, bucket.lr_mult, bucket.decay_mult)]
            return
        for p, h, lm, dm in zip(bucket.params, bucket.history, bucket.lr_mult, bucket.decay_mult):
            g = grads.get(id(p), p.grad)
            if self.reduce == "mean":
                g = g / ws
            apply_rule(hy, p.data, g, h, lm, dm, decay_scale=float(ws) if self.reduce == "sum" else 1.0)

    def finish_iteration(self):
        if not (self.deferred_step and self._deferred):
            return
        ws = self.sync.rank_ctx.world_size
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)      # every bucket's all-reduce has been issued there
        foreach_sgd_step(self.sync.hyper, self._deferred, self.lr_t,
                         decay_scale=float(ws) if self.reduce == "sum" else 1.0,
                         gscale=1.0 / ws if self.reduce == "mean" else 1.0)
        self._deferred = []

    def launch(self, bucket):
        if not self.cuda:
            self._reduce_and_step(bucket)
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for p in bucket.params:
                p.grad.record_stream(self.stream)
            self._reduce_and_step(bucket)
            if bucket.event is None:
                bucket.event = torch.cuda.Event()
            bucket.event.record(self.stream)

    def bytes_on_wire(self):
        out = {"dense_allreduce_bytes": self.dense_bytes}
        st = getattr(self.sync, "sfb_stats", None)          # library sufficient-factor path (parallel/sfb.py)
        if st is not None:
            out["sfb_bytes"], out["sfb_dense_equiv_bytes"] = st.sfb_bytes, st.dense_equiv_bytes
        return out


def _apply_global_grad_step(sync, hyper, p, h, lr_mult, decay_mult):
    """Step a parameter whose gradient is already the global sum (library SFB): equals the sum of the per-worker
    updates the SSP backends would otherwise exchange (decay × world size, like TorchDistBackend's `sum` mode)."""
    ws = sync.rank_ctx.world_size
    delta = compute_update(hyper, p.data, p.grad, h, lr_mult, decay_mult, decay_scale=float(ws))
    with torch.no_grad():
        p.data.sub_(delta)
    sync.invalidate_operands(p)


class SSPBackend(Backend):
    """Stale-synchronous-parallel async SGD without a server.

    Worker p at clock c computes its own step Δ_p(c) (own momentum history, as in the
    reference), applies it locally at once (read-my-writes) and starts an asynchronous
    all-reduce of Δ(c).  The summed remote part Σ_{q≠p} Δ_q(c) is folded into the weights as
    soon as the collective finishes, and *must* have been folded in before clock c+s+1 starts —
    Bösen's guarantee that a reader at clock c sees all updates with clock ≤ c−s−1.
    ``staleness=0`` degenerates to BSP with summed per-worker updates (SSPPush, s=0).

    reference: ps/src/petuum_ps/consistency/ssp_consistency_controller.cpp:37-161,
    ssp_push_consistency_controller.cpp:70-115, SURVEY Appendix B "Consistency contract"."""
    name = "ssp"

    def __init__(self, staleness=0, delay_hook: Optional[Callable[[int], None]] = None):
        self.staleness = int(staleness)
        self.inflight: deque = deque()     # (clock, [(param, delta_sum, own, work)])
        self.clock = 0
        self.delay_hook = delay_hook       # fault/delay injection for tests
        self.max_observed_lag = 0
        self.wire_bytes = 0

    def setup(self, sync):
        super().setup(sync)
        self.cur: List = []

    def launch(self, bucket):
        hy = self.sync.hyper
        for p, h, lm, dm in zip(bucket.params, bucket.history, bucket.lr_mult, bucket.decay_mult):
            if getattr(p, "_grad_is_global", False):
                # sufficient-factor weights: .grad is already Σ_p u_pᵀv_p on every rank (parallel/sfb.py), so the summed
                # update Σ_p Δ_p is one local step on it with the decay counted once per worker — nothing to exchange
                _apply_global_grad_step(self.sync, hy, p, h, lm, dm)
                continue
            delta = compute_update(hy, p.data, p.grad, h, lm, dm)
            with torch.no_grad():
                p.data.sub_(delta)                       # read-my-writes
            # the wire buffer must be dense (NCCL rejects channels-last strided conv weights): logical-order copy
            total = delta.contiguous().clone() if not delta.is_contiguous() else delta.clone()
            work = dist.all_reduce(total, async_op=True) if self.sync.rank_ctx.distributed else None
            self.wire_bytes += total.numel() * total.element_size()
            self.cur.append((p, total, delta, work))

    def _fold(self, entry):
        for p, total, own, work in entry:
            if work is not None:
                work.wait()
            with torch.no_grad():
                p.data.sub_(total - own)
            self.sync.invalidate_operands(p)         # bf16 shadows of the sm100 engine follow the fp32 master

    def finish_iteration(self):
        if self.delay_hook is not None:
            self.delay_hook(self.clock)
        self.inflight.append((self.clock, self.cur))
        self.cur = []
        self.clock += 1
        # fold in everything that has already completed (opportunistic freshness) ...
        while self.inflight:
            clk, entry = self.inflight[0]
            done = all(w is None or w.is_completed() for (_, _, _, w) in entry)
            must = clk <= self.clock - 1 - self.staleness
            if not (done or must):
                break
            self.max_observed_lag = max(self.max_observed_lag, self.clock - 1 - clk)
            self._fold(entry)
            self.inflight.popleft()

    def drain(self):
        while self.inflight:
            _, entry = self.inflight.popleft()
            self._fold(entry)

    def bytes_on_wire(self):
        return {"ssp_delta_bytes": self.wire_bytes}


class SSPAggrBackend(Backend):
    """Bösen's SSPAggr consistency model: bandwidth-managed, magnitude-prioritised communication under a staleness bound.

    Every clock a worker computes its own step Δ (own momentum history), applies it locally (read-my-writes) and adds it
    to a per-parameter *residual* of not-yet-communicated updates.  Only the ``fraction`` of residual entries with the
    largest relative magnitude |r| / (|w| + ε) — the reference's ``update_sort_policy = RelativeMagnitude`` under a
    ``client_bandwidth_mbps`` budget — is exchanged (index / value pairs, all-gather) and folded into the peers' weights;
    the rest waits.  Every ``staleness + 1`` clocks the whole residual is flushed, so no update is ever older than the
    bound (the reference flushes its oplogs on the clock message for the same reason), and each residual entry is
    sent exactly once.  ``fraction = 1`` degenerates to BSP with summed per-worker updates.

    reference: ps/src/petuum_ps/thread/ssp_aggr_bg_worker.cpp:576-647 (budgeted oplog sends),
    ps/src/petuum_ps/server/server_table.cpp:263-287 + ps/src/petuum_ps_common/storage/numeric_container_row.hpp:21-29
    (importance = relative magnitude), ssp_aggr_server_thread.cpp:13-55."""
    name = "ssp_aggr"

    def __init__(self, staleness=0, fraction=0.1, eps=1e-8):
        self.staleness = int(staleness)
        self.fraction = float(fraction)
        if not 0.0 < self.fraction <= 1.0:
            raise ValueError("SSPAggr: fraction of the update sent per clock must be in (0, 1]")
        self.eps = eps
        self.clock = 0
        self.wire_bytes = 0
        self.dense_equiv_bytes = 0
        self.residual = {}            # id(param) -> flat fp32 residual
        self.touched = []

    def launch(self, bucket):
        hy = self.sync.hyper
        for p, h, lm, dm in zip(bucket.params, bucket.history, bucket.lr_mult, bucket.decay_mult):
            if getattr(p, "_grad_is_global", False):
                _apply_global_grad_step(self.sync, hy, p, h, lm, dm)      # see SSPBackend.launch
                continue
            delta = compute_update(hy, p.data, p.grad, h, lm, dm)
            with torch.no_grad():
                p.data.sub_(delta)                                   # read-my-writes
                r = self.residual.get(id(p))
                if r is None:
                    r = self.residual[id(p)] = torch.zeros(p.numel(), dtype=torch.float32, device=p.device)
                r.add_(delta.reshape(-1).float() if delta.is_contiguous() else delta.contiguous().reshape(-1).float())
            self.touched.append(p)

    def _exchange(self, p, flush: bool):
        r = self.residual[id(p)]
        n = r.numel()
        k = n if flush else max(1, int(math.ceil(self.fraction * n)))
        rc = self.sync.rank_ctx
        with torch.no_grad():
            if k >= n:
                idx = torch.arange(n, device=r.device)
            else:
                w = p.data.contiguous().reshape(-1).float() if not p.data.is_contiguous() else p.data.reshape(-1).float()
                idx = torch.topk(r.abs() / (w.abs() + self.eps), k, sorted=False).indices
            val = r[idx]
            r[idx] = 0.0                                             # sent exactly once
            self.wire_bytes += k * (4 + (0 if k >= n else 4))
            self.dense_equiv_bytes += n * 4
            if not rc.distributed:
                return
            ws = rc.world_size
            vals = [torch.empty_like(val) for _ in range(ws)]
            dist.all_gather(vals, val)
            if k >= n:
                total = torch.stack(vals).sum(0) - val               # everybody else's flushed residual (logical order)
                p.data.sub_(total.view(p.shape))
            else:
                idxs = [torch.empty_like(idx) for _ in range(ws)]
                dist.all_gather(idxs, idx)
                upd = torch.zeros(n, dtype=torch.float32, device=r.device)
                for q in range(ws):
                    if q != rc.rank:
                        upd.index_add_(0, idxs[q], vals[q])
                p.data.sub_(upd.view(p.shape))
        self.sync.invalidate_operands(p)

    def finish_iteration(self):
        self.clock += 1
        flush = self.fraction >= 1.0 or (self.clock % (self.staleness + 1) == 0)
        seen = set()
        for p in self.touched:
            if id(p) in seen:
                continue
            seen.add(id(p))
            self._exchange(p, flush)
        self.touched = []

    def drain(self):
        """Flush every residual (end of training / before a snapshot or a test pass)."""
        for b in self.sync.buckets:
            for p in b.params:
                if id(p) in self.residual:
                    self._exchange(p, True)

    def bytes_on_wire(self):
        return {"ssp_aggr_bytes": self.wire_bytes, "dense_equiv_bytes": self.dense_equiv_bytes}


class GradSync:
    """Hooks autograd to the backend (DWBP) and tracks per-bucket completion events."""

    def __init__(self, net, rank_ctx, hyper: Hyper, backend: Backend):
        self.net, self.rank_ctx, self.hyper, self.backend = net, rank_ctx, hyper, backend
        self.buckets = build_buckets(net)
        self.bucket_of: Dict[int, Bucket] = {}
        self._handles = []
        self.launch_order: List[int] = []
        for b in self.buckets:
            for p in b.params:
                self.bucket_of[id(p)] = b
        backend.setup(self)
        self.attach()

    def attach(self):
        for b in self.buckets:
            for p in b.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        if self.backend.uses_comm_stream:
            for b in self.buckets:
                self._handles.append(b.layer.register_forward_pre_hook(self._make_wait(b)))

    def detach(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def _make_wait(self, bucket):
        def wait(module, args):
            if bucket.event is not None:
                torch.cuda.current_stream().wait_event(bucket.event)
            if bucket.sfb_event is not None:
                torch.cuda.current_stream().wait_event(bucket.sfb_event)
        return wait

    def begin_iteration(self, lr: float):
        self.hyper.lr = lr
        if hasattr(self.backend, "set_lr"):
            self.backend.set_lr(lr)
        self.launch_order = []
        for b in self.buckets:
            b.pending = len(b.params) - len(b.self_updating)

    def _on_grad(self, p):
        b = self.bucket_of[id(p)]
        # autograd also fires this hook when a Function returned None for the parameter (weights that were
        # stepped inside their layer's backward by the fused SFB / wgrad kernel): those do not count
        if p.grad is None or id(p) in b.self_updating:
            return
        b.pending -= 1
        if b.pending == 0:
            self.launch_order.append(b.id)
            with nvtx_range(f"sync/{getattr(b.layer, 'layer_name', b.id)}"):
                self.backend.launch(b)
            if not getattr(self.backend, "manages_operands", False):
                st = getattr(b.layer, "_sm100", None)
                if st is not None:
                    st.mark_updated()                # library backends step the fp32 master only
            for q in b.params:
                q.grad = None

    def invalidate_operands(self, p):
        """A parameter's fp32 master was modified outside the fused kernels: its layer's bf16 operands are stale."""
        b = self.bucket_of.get(id(p))
        st = getattr(b.layer, "_sm100", None) if b is not None else None
        if st is not None:
            st.mark_updated()

    def weights_changed(self):
        """The fp32 masters were rewritten wholesale (``.caffemodel`` load / restore, followed by the broadcast from rank
        0): every rank re-derives its bf16 operands — including the arena-resident shadows that are otherwise only
        written by the update kernels."""
        for layer in self.net.layers:
            st = getattr(layer, "_sm100", None)
            if st is not None:
                st.mark_updated()
        if hasattr(self.backend, "refresh_shadows"):
            self.backend.refresh_shadows()

    def finish_iteration(self):
        """Launch buckets whose grads never materialised this step (unused layers) — none in
        practice — and let the backend close the clock.  Does *not* host-sync."""
        self.backend.finish_iteration()

    def wait_all(self):
        """JoinSyncThreads equivalent (stream-level, no host sync)."""
        if self.backend.uses_comm_stream:
            cur = torch.cuda.current_stream()
            for b in self.buckets:
                if b.event is not None:
                    cur.wait_event(b.event)
                if b.sfb_event is not None:
                    cur.wait_event(b.sfb_event)

    # ---- optimizer state for .solverstate ------------------------------------------------
    def history_tensors(self) -> List[torch.Tensor]:
        out = []
        by_param = {}
        for b in self.buckets:
            for p, h in zip(b.params, b.history):
                by_param[id(p)] = h
        for p in self.net.params:
            out.append(by_param.get(id(p), torch.zeros_like(p)))
        return out

    def load_history(self, tensors: List[torch.Tensor]):
        hs = self.history_tensors()
        if len(tensors) != len(hs):
            raise ValueError("Incorrect length of history blobs.")
        for h, t in zip(hs, tensors):
            with torch.no_grad():
                h.copy_(t.reshape(h.shape))
