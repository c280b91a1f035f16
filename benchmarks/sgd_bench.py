"""Fused outer-product + optimizer-step GEMM (EPI_SGD) micro-benchmark: fc6 / fc7 shapes of AlexNet.
    python benchmarks/sgd_bench.py [iters]
Reports time and achieved HBM traffic (18 B per weight: W, H read + written, bf16 shadow written)."""
import sys
import torch
sys.path.insert(0, ".")
from poseidon_b200.ops import sm100

K = sm100.K()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for (M, N, Kd) in ((256, 4096, 9216), (256, 4096, 4096)):
    dy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    x = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, Kd, device="cuda")
    h = torch.zeros_like(w)
    wb = w.to(torch.bfloat16)
    lr_t = torch.ones(1, device="cuda")
    def run():
        K.sfb_outer_sgd([dy.data_ptr()], [x.data_ptr()], M, N, Kd, w, h, wb, 1.0, 0.01, 0.9, 5e-4, 0, False, 1e-8,
                        None, 0, 0, 0, 0, lr_t, None)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"sgd-gemm {N}x{Kd} (batch {M}): {ms*1e3:8.1f} us   {N*Kd*18/ms/1e6:7.0f} GB/s of 18 B/weight traffic")
    # reference points: the unfused path = dense wgrad GEMM (fp32 out) + the elementwise optimizer kernel
    g = torch.empty_like(w)
    def unfused():
        K.gemm_f32(dy, True, x, True, g, 1.0, False, 1, 0)
        K.fused_update(w, g, h, wb, 0.01, 0.9, 5e-4, 0, False, 1e-8, 1.0, lr_t)
    for _ in range(2):
        unfused()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        unfused()
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"   unfused gemm_f32 + fused_update     : {ms2*1e3:8.1f} us   ({N*Kd*30/ms2/1e6:7.0f} GB/s of 30 B/weight traffic)")
