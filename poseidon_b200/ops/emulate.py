"""CPU emulation of ``torch.ops.poseidon`` — the same op surface, the same operand layouts, plain PyTorch math.

Why: the sm100 engine is two things — the CUDA kernels in ``csrc/`` and a fair amount of Python around them that decides
operand layouts (ROW / space-to-depth / 8-channel first layers, channel-padded K, (h, w, c) inner-product weights), keeps
bf16 shadows in step with fp32 masters, plans epilogue fusion and routes gradients into the optimizer kernels.  The kernels
can only be tested on a B200; the Python around them is just as easy to break, and with ``POSEIDON_EMULATE=1`` it runs on
any machine: ``sm100.K()`` hands out this module's :class:`EmulatedKernels` instead of the compiled extension, and the
whole engine (``Solver(engine="sm100")``, fusion plan, FusedBackend's single-process path) steps on the CPU.

Every method documents the layout contract of the kernel it stands in for (file:line of the CUDA implementation); the
arithmetic is fp32 with one bf16 rounding at the output, which is also what the kernels do (fp32 accumulation in TMEM,
bf16 store), so results agree with the GPU to bf16 rounding.  This is a *test double*: it is never selected implicitly,
it is slow, and the multi-GPU peer-memory ops (which take raw device pointers) are not provided.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import reference as R

CL = torch.channels_last
BF16 = torch.bfloat16


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous(memory_format=CL)


def _storage_order_flat(t: torch.Tensor) -> torch.Tensor:
    """Elements of a dense tensor in memory order."""
    dims = sorted(range(t.dim()), key=lambda d: (t.shape[d] == 1, -t.stride(d)))
    return t.permute(dims).reshape(-1)


def _act(v: torch.Tensor, relu: bool, slope: float) -> torch.Tensor:
    if not relu:
        return v
    return torch.where(v > 0, v, v * slope)


class EmulatedKernels:
    """Drop-in for ``torch.ops.poseidon`` (single-process ops)."""

    emulated = True

    # ------------------------------------------------------------------------------------------ switches
    def set_conv_cluster(self, c):           # csrc/gemm/conv_ops.cu (tuning knobs: no numerical effect)
        return None

    def set_conv_im2col(self, on):
        return None

    # ------------------------------------------------------------------------------------------ GEMM
    def gemm_bf16(self, a, a_mn, b, b_mn, bias, relu, slope, mask, out, bn):
        """C[M,N] bf16 = act(A·Bᵀ + bias), A = a ([M,K]) or aᵀ, B = b ([N,K]) or bᵀ (csrc/gemm/gemm_ops.cu:173-240)."""
        assert a.dtype == BF16 and b.dtype == BF16
        A = a.t() if a_mn else a
        B = b.t() if b_mn else b
        v = A.float() @ B.float().t()
        if bias is not None:
            v = v + bias.float()
        v = _act(v, relu, slope)
        if mask is not None:
            v = torch.where(mask.float() > 0, v, v * slope)
        if out is None:
            return v.to(BF16)
        out.copy_(v)
        return out

    def gemm_f32(self, a, a_mn, b, b_mn, out, alpha, accumulate, split_k, bn):
        """out[M,N] fp32 (+)= alpha · A·Bᵀ (csrc/gemm/gemm_ops.cu:243-300)."""
        A = a.t() if a_mn else a
        B = b.t() if b_mn else b
        v = (A.float() @ B.float().t()) * alpha
        if accumulate:
            out.add_(v)
        else:
            out.copy_(v)

    # ------------------------------------------------------------------------------------------ convolution
    @staticmethod
    def _w4_tap(wb, cout, r, s, cg, cgk):
        """[Cout, R*S*Cgk] (K ordered (r, s, c), slots >= Cg zero) -> [Cout, Cg, R, S] fp32."""
        cgk = cgk or cg
        return wb.reshape(cout, r, s, cgk)[..., :cg].permute(0, 3, 1, 2).float()

    def conv_fprop(self, x, wb, bias, kernel, stride, pad, groups, mode, OH, OW, relu, slope, out):
        """Implicit-GEMM forward (csrc/gemm/conv_ops.cu:176-247).
        mode 0 (TAP): x logical [N,C,H,W] channels-last; wb [Cout, R*S*Cgk].
        mode 1 (ROW): x is the spatially pre-padded NHWC4 image; wb [Cout, R*Lp], each kernel row = S*4 contiguous
        elements padded to Lp = roundup8(S*4)."""
        r, s = kernel
        cout = wb.shape[0]
        if mode == 1:
            cp = x.shape[1]
            lp = wb.shape[1] // r
            w4 = wb.reshape(cout, r, lp)[:, :, : s * cp].reshape(cout, r, s, cp).permute(0, 3, 1, 2).float()
            y = F.conv2d(x.float(), w4, None, tuple(stride), 0)[:, :, :OH, :OW]
        else:
            cg = x.shape[1] // groups
            cgk = wb.shape[1] // (r * s)
            w4 = self._w4_tap(wb, cout, r, s, cg, cgk)
            y = F.conv2d(x.float(), w4, None, tuple(stride), tuple(pad), 1, groups)
        assert tuple(y.shape[2:]) == (OH, OW), (tuple(y.shape), OH, OW)
        if bias is not None:
            y = y + bias.float().view(1, -1, 1, 1)
        y = _act(y, relu, slope)
        if out is not None:
            out.copy_(y)
            return out
        return _nhwc(y.to(BF16))

    def conv_dgrad(self, dy, wt, kernel, pad, groups, H, W, mask, slope):
        """Stride-1 data gradient from the packed operand wt [groups*Cg, R*S*Cop] (K ordered (r, s, co));
        optional producer-ReLU mask applied in the epilogue (csrc/gemm/conv_ops.cu:250-301)."""
        r, s = kernel
        cout = dy.shape[1]
        cout_g = cout // groups
        cg = wt.shape[0] // groups
        cop = wt.shape[1] // (r * s)
        w = wt.reshape(groups, cg, r, s, cop)[..., :cout_g].float()           # g, ci, r, s, co
        w4 = w.permute(0, 4, 1, 2, 3).reshape(cout, cg, r, s)                  # conv weight [Cout, Cg, R, S]
        dx = F.conv_transpose2d(dy.float(), w4, None, 1, tuple(pad), 0, groups)
        assert tuple(dx.shape[2:]) == (H, W)
        if mask is not None:
            dx = torch.where(mask.float() > 0, dx, dx * slope)
        return _nhwc(dx.to(BF16))

    def conv_wgrad(self, x, dy, dw, kernel, stride, pad, groups, mode, alpha, cgk):
        """dw[Cout, K] fp32 += alpha · (weight gradient in the operand's K order) (csrc/gemm/conv_ops.cu:304-365)."""
        r, s = kernel
        cout = dy.shape[1]
        oh, ow = dy.shape[2:]
        if mode == 1:
            cp = x.shape[1]
            lp = dw.shape[1] // r
            xc = x[:, :, : (oh - 1) * stride[0] + r, : (ow - 1) * stride[1] + s].float()
            g = torch.nn.grad.conv2d_weight(xc, (cout, cp, r, s), dy.float(), tuple(stride), 0)
            g = g.permute(0, 2, 3, 1).reshape(cout, r, s * cp)
            g = F.pad(g, (0, lp - s * cp)).reshape(cout, r * lp)
        else:
            cg = x.shape[1] // groups
            g = torch.nn.grad.conv2d_weight(x.float(), (cout, cg, r, s), dy.float(), tuple(stride), tuple(pad), 1, groups)
            g = g.permute(0, 2, 3, 1).reshape(cout, r * s * cg)
        dw.add_(g, alpha=alpha)

    def conv_pack_dgrad(self, w, Cout, RS, Cg, groups, out, cop):
        """fp32 [Cout][RS][Cg] -> bf16 wt [groups*Cg, RS*Cop], co padded with zeros (csrc/gemm/conv_ops.cu:366-398)."""
        cout_g = Cout // groups
        cop = cop if cop > 0 else cout_g
        t = w.reshape(groups, cout_g, RS, Cg).permute(0, 3, 2, 1)              # g, ci, tap, co
        t = F.pad(t, (0, cop - cout_g)).reshape(groups * Cg, RS * cop)
        if out is None:
            return t.to(BF16).contiguous()
        out.copy_(t.reshape(out.shape))
        return out

    def conv_pack_padded(self, wb, Cout, RS, Cg, Cgk, out):
        """bf16 [Cout][RS][Cg] -> [Cout, RS*Cgk], slots >= Cg zero (csrc/gemm/conv_ops.cu:400-425)."""
        t = F.pad(wb.reshape(Cout, RS, Cg), (0, Cgk - Cg)).reshape(Cout, RS * Cgk)
        if out is None:
            return t.contiguous()
        out.copy_(t)
        return out

    # ------------------------------------------------------------------------------------------ elementwise
    def relu_fwd(self, x, slope):
        return _act(x.float(), True, slope).to(BF16)

    def relu_bwd(self, y, dy, slope):
        assert y.stride() == dy.stride(), "relu_bwd: layout mismatch"
        return torch.where(y.float() > 0, dy.float(), dy.float() * slope).to(BF16)

    def dropout_apply(self, x, ratio, seed, seed_dev):
        """The same (seed, iteration, element) -> keep map for forward and backward (csrc/ops/elementwise.cu:178-222);
        the emulation draws the map from a generator keyed the same way instead of the kernel's integer hash."""
        it = int(seed_dev.item()) if seed_dev is not None else 0
        g = torch.Generator().manual_seed((int(seed) ^ (it * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF)
        keep = torch.rand(x.numel(), generator=g) > ratio
        flat = _storage_order_flat(x).float() * keep * (1.0 / (1.0 - ratio))
        y = torch.empty_like(x)                                   # dense input: same strides
        torch.as_strided(y, (x.numel(),), (1,)).copy_(flat)       # the map is indexed by memory position, like the kernel's
        return y

    def colsum(self, dy, rows, C, ld, out, alpha, accumulate):
        """out[c] (+)= alpha · Σ_rows dy[row*ld + c] over the tensor's memory (csrc/ops/elementwise.cu:224-268)."""
        v = torch.as_strided(dy, (rows, C), (ld, 1), dy.storage_offset()).float().sum(0) * alpha
        if accumulate:
            out.add_(v)
        else:
            out.copy_(v)

    def transform_nhwc(self, x, h_off, w_off, flip, mean, scale, OH, OW, cp, opad, wextra, hextra, s2d):
        """Crop / mirror / mean / scale / channel-pad / spatial pre-pad (/ space-to-depth by 4) in one pass
        (csrc/ops/elementwise.cu:19-148).  Output: bf16 channels-last [N, cp, OH+2opad+hextra, OW+2opad+wextra],
        or [N, 64, ·/4, ·/4] with channel (y%4)*16 + (x%4)*4 + c when s2d."""
        n, c, h, w = x.shape
        xf = x.float()
        if mean is not None:
            xf = xf - (mean.view(1, c, 1, 1) if mean.numel() == c else mean.view(1, c, h, w))
        outs = []
        for i in range(n):
            ho, wo = int(h_off[i]), int(w_off[i])
            t = xf[i, :, ho: ho + OH, wo: wo + OW]
            if int(flip[i]):
                t = t.flip(-1)
            outs.append(t)
        t = torch.stack(outs) * scale
        t = F.pad(t, (opad, opad + wextra, opad, opad + hextra, 0, cp - c))
        if s2d:
            hq, wq = t.shape[2] // 4, t.shape[3] // 4
            t = t.view(n, cp, hq, 4, wq, 4).permute(0, 2, 4, 3, 5, 1).reshape(n, hq, wq, 64)
            return t.to(BF16).permute(0, 3, 1, 2)
        return _nhwc(t.to(BF16))

    # ------------------------------------------------------------------------------------------ LRN / pooling
    def lrn_fwd(self, x, size, alpha, beta, fuse_relu):
        xf = x.float()
        if fuse_relu:
            xf = xf.clamp_min(0)
        return _nhwc(R.lrn_across(xf, size, alpha, beta).to(BF16))

    def lrn_bwd(self, x, dy, size, alpha, beta, fuse_relu):
        """dx of across-channel LRN; fuse_relu zeroes the gradient where the (post-ReLU) input is not positive
        (csrc/ops/lrn.cu:140-210)."""
        xf = x.float().detach().requires_grad_(True)
        with torch.enable_grad():
            y = R.lrn_across(xf, size, alpha, beta)
        (dx,) = torch.autograd.grad(y, xf, dy.float())
        if fuse_relu:
            dx = dx * (x.float() > 0)
        return _nhwc(dx.to(BF16))

    def pool_fwd(self, x, is_max, k, s, p, oh, ow, want_idx):
        """(y, idx).  The kernel's idx is the in-window tap of the maximum, one byte per output element
        (csrc/ops/pool.cu:28-110); it is opaque to the caller, so the emulation stores the plane index instead."""
        xf = x.float()
        if is_max:
            y, idx = F.max_pool2d(xf, tuple(k), tuple(s), tuple(p), ceil_mode=True, return_indices=True)
            y, idx = y[:, :, :oh, :ow], idx[:, :, :oh, :ow]
            assert tuple(y.shape[2:]) == (oh, ow)
            return _nhwc(y.to(BF16)), (idx.contiguous() if want_idx else torch.empty(0, dtype=torch.int64))
        y = R.ave_pool(xf, tuple(k), tuple(s), tuple(p))[:, :, :oh, :ow]
        return _nhwc(y.to(BF16)), torch.empty(0, dtype=torch.int64)

    def pool_bwd(self, dy, idx, is_max, in_hw, k, s, p):
        n, c = dy.shape[:2]
        h, w = in_hw
        if is_max:
            dx = torch.zeros(n, c, h * w, dtype=torch.float32)
            dx.scatter_add_(2, idx.reshape(n, c, -1), dy.float().reshape(n, c, -1))
            return _nhwc(dx.view(n, c, h, w).to(BF16))
        xf = torch.zeros(n, c, h, w, requires_grad=True)
        with torch.enable_grad():
            y = R.ave_pool(xf, tuple(k), tuple(s), tuple(p))
        (dx,) = torch.autograd.grad(y, xf, dy.float())
        return _nhwc(dx.to(BF16))

    # ------------------------------------------------------------------------------------------ loss
    def softmax_xent(self, x, label, grad_scale, want_grad, want_prob):
        """(loss[1] fp32, dx like x, prob fp32): mean NLL over rows, dx = (p - onehot)·grad_scale/rows
        (csrc/ops/softmax_xent.cu:28-107)."""
        rows = x.shape[0]
        prob = torch.softmax(x.float(), 1)
        lab = label.long()
        picked = prob.gather(1, lab.view(-1, 1)).clamp_min(torch.finfo(torch.float32).tiny)
        loss = (-picked.log().sum() / rows).reshape(1)
        dx = torch.empty(0, dtype=x.dtype)
        if want_grad:
            g = prob.clone()
            g.scatter_add_(1, lab.view(-1, 1), torch.full((rows, 1), -1.0))
            dx = (g * (grad_scale / rows)).to(x.dtype)
        return loss, dx, (prob if want_prob else torch.empty(0))

    # ------------------------------------------------------------------------------------------ optimizer
    def fused_update(self, w, g, h, wb, lr, momentum, decay, rule, l1, delta, gscale, lr_dev):
        """One optimizer step in place on (w, h) + refresh of the bf16 shadow in the master's storage order
        (csrc/comm/fused_update.cu:30-108).  rule 0 SGD, 1 Nesterov, 2 AdaGrad."""
        for t in (g, h):
            assert t.shape == w.shape and all(t.stride(d) == w.stride(d) for d in range(w.dim()) if w.shape[d] > 1), \
                "fused_update: layout mismatch"
        if lr_dev is not None:
            lr = lr * float(lr_dev[0])
        g = g.float() * gscale
        if decay != 0:
            g = g + decay * (torch.sign(w) if l1 else w)
        if rule == 0:
            h.mul_(momentum).add_(g, alpha=lr)
            w.sub_(h)
        elif rule == 1:
            h_old = h.clone()
            h.mul_(momentum).add_(g, alpha=lr)
            w.sub_((1.0 + momentum) * h - momentum * h_old)
        else:
            h.add_(g * g)
            w.sub_(lr * g / (h.sqrt() + delta))
        if wb is not None:
            wb.view(-1).copy_(_storage_order_flat(w))

    # ------------------------------------------------------------------------------------------ multi-GPU (not emulated)
    def _peer(self, *a, **k):
        raise NotImplementedError("peer-memory ops take raw device pointers and are not emulated on the CPU")

    allreduce_sgd = peer_push = peer_signal = sfb_outer_sgd = sfb_outer_f32 = _peer
