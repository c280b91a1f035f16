"""Opt-in validation of the EXPERIMENTAL paired-CTA (cta_group::2) GEMM in csrc_experimental/ (DESIGN.md §8).

    POSEIDON_EXPERIMENTAL=1 python -m pytest tests/test_pair_gemm_gpu.py -q

Skipped otherwise: the kernel compiled and its SASS shows UTCHMMA.2CTA / UTMALDG.2D.2CTA / UTCBAR.2CTA.MULTICAST, but it has
not run on hardware yet."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("POSEIDON_EXPERIMENTAL") != "1", reason="experimental kernel: opt-in")]


@pytest.fixture(scope="module")
def exp():
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from poseidon_b200.ops import build
    build.load_experimental()
    return torch.ops.poseidon_exp


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (4096, 4096, 1024), (384, 256, 512), (1000, 320, 200)])
def test_pair_gemm_matches_fp32_reference(exp, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
    c = exp.pair_gemm_bf16(a, b)
    ref = a.float() @ b.float().t()
    err = (c.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, err


def test_pair_gemm_throughput(exp):
    a = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
    b = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        exp.pair_gemm_bf16(a, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        exp.pair_gemm_bf16(a, b)
    e1.record()
    torch.cuda.synchronize()
    tf = 2 * 8192 ** 3 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print(f"pair_gemm 8192^3: {tf:.0f} TFLOP/s (single-CTA kernel: 1317, cuBLAS: 1647)")
    assert tf > 200
