import sys
import torch
sys.path.insert(0, ".")
from poseidon_b200.ops import sm100, reference as R
K = sm100.K()
torch.manual_seed(0)
N, C, H, W, size, alpha, beta = 2, 32, 12, 10, 3, 0.5, 0.75
x = (torch.randn(N, C, H, W, device="cuda") * 2).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
dy = torch.randn(N, C, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
xr = x.float().requires_grad_(True)
yr = R.lrn_within(xr, size, alpha, beta)
yr.backward(dy.float())
y = K.lrn_within_fwd(x, size, alpha, beta)
dx = K.lrn_within_bwd(x, dy, size, alpha, beta)
print("fwd err", (y.float() - yr).abs().max().item())
e = (dx.float() - xr.grad).abs()
print("bwd err", e.max().item(), "of", xr.grad.abs().max().item())
idx = (e == e.max()).nonzero()[0].tolist()
print("worst at (n,c,h,w) =", idx, "got", dx[tuple(idx)].item(), "ref", xr.grad[tuple(idx)].item())
print("err by h:", e.amax((0, 1, 3)).tolist())
print("err by w:", e.amax((0, 1, 2)).tolist())
print("err by c:", e.amax((0, 2, 3)).tolist())
print("err by n:", e.amax((1, 2, 3)).tolist())

print("---- special cases")
for (sz, al) in ((3, 0.0), (1, 0.5), (3, 0.5)):
    xr = x.float().requires_grad_(True)
    yr = R.lrn_within(xr, sz, al, beta)
    yr.backward(dy.float())
    dxk = K.lrn_within_bwd(x, dy, sz, al, beta).float()
    # torch replica of the kernel's two passes
    xf, g = x.float(), dy.float()
    pre = (sz - 1) // 2
    ssum = torch.nn.functional.avg_pool2d(xf * xf, sz, 1, pre, count_include_pad=True) * (sz * sz)
    scale = 1 + al * ssum / (sz * sz)
    p = scale ** (-beta)
    r = g * xf * p / scale / (sz * sz)
    rs = torch.nn.functional.avg_pool2d(r, sz, 1, pre, count_include_pad=True) * (sz * sz)
    dxt = g * p - 2 * al * beta * xf * rs
    print(f"size {sz} alpha {al}: kernel-vs-autograd {(dxk - xr.grad).abs().max().item():.4f}  replica-vs-autograd "
          f"{(dxt - xr.grad).abs().max().item():.6f}  kernel-vs-(dy*p) {(dxk - g * p).abs().max().item():.4f}")
